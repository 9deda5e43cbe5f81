"""Worker of tests/test_sharded_gpu.py: one rank of a world that reads ONE Arrow expression table through shard_source +
SLAFIterableDataset on a GPU (rank % visible GPUs) and reports the cells it emitted; torch.distributed (gloo) carries the
every-cell-once reduction.  Launched by torch.distributed.run."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from oracle import c_oracle
from slaf_b200 import synthetic
from slaf_b200.distributed import shard_rows, shard_source
from slaf_b200.ml import GeneformerTokenizer, SLAFIterableDataset
from slaf_b200.ml.sources import ArrowExpressionSource, SyntheticSLAFArray, write_arrow_dataset


def main():
    path, route = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    n_cells, n_genes = 1500, 30000
    b = synthetic.make_batch(n_cells, n_genes, 150, seed=31, min_nnz=5)
    if rank == 0:
        write_arrow_dataset(os.path.join(path, "expression.arrow"), b.cell, b.gene, b.value, max_rows_per_file=20011, rows_per_batch=4096)
    dist.barrier()
    src = ArrowExpressionSource(os.path.join(path, "expression.arrow"))
    csi = np.asarray(b.cell_start, np.int64)
    rows = [f.count_rows() for f in src.get_fragments()]
    lo, hi = shard_rows(rows, csi, world, rank)
    arr = SyntheticSLAFArray(shard_source(src, csi, world, rank), csi, n_genes)
    tok = GeneformerTokenizer(arr, max_genes=256, device=dev)
    loader = SLAFIterableDataset(arr, tok, batch_size=64, use_mixture_of_scanners=False, by_fragment=True, verbose=False, device=dev,
                                 use_cell_start_index={"csi": None, "coo": False}[route])
    ids_all, ok = [], True
    want = c_oracle.tokenize(b.cell, b.gene, b.value, "geneformer", 256)
    row_of = {int(c): i for i, c in enumerate(want[3])}
    for batch in loader:
        ids = batch["cell_ids"].cpu().numpy()
        tok_ids, mask = batch["input_ids"].cpu().numpy(), batch["attention_mask"].cpu().numpy()
        for r, c in enumerate(ids):
            i = row_of[int(c)]
            ok = ok and np.array_equal(tok_ids[r], want[0][i]) and np.array_equal(mask[r], want[1][i])
        ids_all.extend(int(c) for c in ids)
    used_csi = loader.batch_processor._csi_arr is not None
    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "ids": ids_all, "rows": [lo, hi], "tokens_ok": bool(ok), "used_csi": used_csi, "device": str(dev)})
    if rank == 0:
        json.dump(gathered, open(os.path.join(path, "result.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
