"""GPU parity tests: the sm_100a kernels, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Bit-exact (integer / byte / index work): no tolerance anywhere."""

import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import c_oracle
from slaf_b200 import _native, synthetic
from tests.util import adversarial_batch, assert_same

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "tokenizer_golden.npz"))
KATS = json.load(open(os.path.join(HERE, "golden", "window_kats.json")))["cases"]

MODES = [("geneformer", None), ("geneformer_pct", 37.5), ("scgpt_raw", None), ("scgpt_binned", None)]


def py_perm(seed, n):
    random.seed(seed)
    idx = list(range(n))
    random.shuffle(idx)
    return idx


def to_dev(a):
    return torch.from_numpy(a).cuda()


def compare(batch, want, mode):
    ids, mask, vals, cids = want
    assert batch.n_cells == len(cids)
    assert batch.input_ids.dtype == torch.int64 and batch.attention_mask.dtype == torch.bool
    assert batch.cell_ids.dtype == torch.int64
    assert_same(batch.cell_ids.cpu().numpy(), cids, "cell_ids")
    assert_same(batch.input_ids.cpu().numpy(), ids, "input_ids")
    assert_same(batch.attention_mask.cpu().numpy(), mask, "attention_mask")
    if mode.startswith("scgpt"):
        assert batch.values.dtype == torch.int64
        assert_same(batch.values.cpu().numpy(), vals, "values")
    else:
        assert batch.values is None and vals is None


@pytest.mark.parametrize("location", ["device", "host"])
@pytest.mark.parametrize("gene_dtype,cell_dtype", [("uint16", "uint32"), ("int32", "int32")])
@pytest.mark.parametrize("value_dtype", ["uint16", "float32"])
@pytest.mark.parametrize("mode,pct", MODES)
def test_adversarial_cells_all_modes(engine, mode, pct, value_dtype, gene_dtype, cell_dtype, location):
    k = 64
    cell, gene, value = adversarial_batch(11, value_dtype, gene_dtype, cell_dtype, k=k)
    C = len(np.unique(cell))
    seed = 42 + 17
    want = c_oracle.tokenize(cell, gene, value, mode, k, perm=py_perm(seed, C), n_bins=51, vocab_size=65000, min_percentile=pct)
    cols = (to_dev(cell), to_dev(gene), to_dev(value)) if location == "device" else (cell, gene, value)
    got = engine.tokenize(*cols, mode=mode, max_genes=k, n_bins=51, vocab_size=65000, min_percentile=pct, shuffle_seed=seed)
    compare(got, want, mode)


@pytest.mark.parametrize("mode,pct", MODES)
@pytest.mark.parametrize("k,max_genes", [(1, 1), (2, 7), (5, 3), (2048, 2048), (1200, 1200), (300, 2048)])
def test_width_and_k_combinations(engine, mode, pct, k, max_genes):
    """max_items != max_genes, tiny widths (CLS only / SEP dropped), odd L (unaligned rows)."""
    cell, gene, value = adversarial_batch(5, "uint16", "uint16", "uint32", k=max(k, 8) if k < 1000 else 40, big=(k >= 300))
    want = c_oracle.tokenize(cell, gene, value, mode, max_genes, max_items=k, n_bins=10, vocab_size=50000, min_percentile=pct)
    got = engine.tokenize(to_dev(cell), to_dev(gene), to_dev(value), mode=mode, max_genes=max_genes, max_items=k,
                          n_bins=10, vocab_size=50000, min_percentile=pct)
    compare(got, want, mode)


@pytest.mark.parametrize("n_bins", [1, 2, 3, 5, 10, 51, 1000])
@pytest.mark.parametrize("value_dtype", ["uint16", "float32"])
def test_binning_all_bin_counts(engine, n_bins, value_dtype):
    cell, gene, value = adversarial_batch(23, value_dtype, "uint16", "uint32", k=32, big=False)
    for mode in ("scgpt_binned", "scgpt_raw"):
        want = c_oracle.tokenize(cell, gene, value, mode, 32, n_bins=n_bins, vocab_size=1000)
        got = engine.tokenize(cell, gene, value, mode=mode, max_genes=32, n_bins=n_bins, vocab_size=1000)
        compare(got, want, mode)


def test_explicit_perm_and_identity(engine):
    cell, gene, value = adversarial_batch(3, "uint16", "uint16", "uint32", k=16, big=False)
    C = len(np.unique(cell))
    perm = np.random.default_rng(0).permutation(C)
    want = c_oracle.tokenize(cell, gene, value, "geneformer", 16, perm=perm)
    got = engine.tokenize(cell, gene, value, mode="geneformer", max_genes=16, perm=perm)
    compare(got, want, "geneformer")
    want = c_oracle.tokenize(cell, gene, value, "geneformer", 16)
    got = engine.tokenize(cell, gene, value, mode="geneformer", max_genes=16)
    compare(got, want, "geneformer")


@pytest.mark.parametrize("cfg", ["cfg1_geneformer", "cfg2_scgpt_binned", "scgpt_raw_default", "float32_geneformer"])
def test_synthetic_configs_full_compare(engine, cfg):
    """BASELINE.json configs 1 and 2 at a size the oracle finishes in seconds (every cell compared)."""
    if cfg == "cfg1_geneformer":
        b = synthetic.make_batch(1500, 20000, 2000, seed=1)
        kw = dict(mode="geneformer", max_genes=2048)
    elif cfg == "cfg2_scgpt_binned":
        b = synthetic.make_batch(1500, 60000, 2000, seed=2)
        kw = dict(mode="scgpt_binned", max_genes=1200, n_bins=51, vocab_size=65000)
    elif cfg == "scgpt_raw_default":
        b = synthetic.make_batch(1500, 60000, 2000, seed=4)
        kw = dict(mode="scgpt_raw", max_genes=1200, n_bins=51, vocab_size=65000)
    else:
        b = synthetic.make_batch(800, 62000, 2500, seed=5, value_dtype="float32")
        kw = dict(mode="geneformer", max_genes=2048)
    seed = 42 + 0
    okw = {("n_bins" if k == "n_bins" else k): v for k, v in kw.items() if k != "mode"}
    want = c_oracle.tokenize(b.cell, b.gene, b.value, kw["mode"], perm=py_perm(seed, b.n_cells), **okw)
    got = engine.tokenize(to_dev(b.cell), to_dev(b.gene), to_dev(b.value), shuffle_seed=seed, **kw)
    compare(got, want, kw["mode"])
    t = engine.timing()
    assert t["kernel_launches"] >= 2


def test_pipelined_submit_collect_two_in_flight(engine):
    """submit(i+1) before collect(i): results must not depend on the overlap."""
    batches = [synthetic.make_batch(300 + 50 * i, 30000, 600, seed=100 + i) for i in range(5)]
    kw = dict(mode="scgpt_binned", max_genes=512, n_bins=51, vocab_size=65000)
    slots = [engine.submit(batches[0].cell, batches[0].gene, batches[0].value)]
    outs = []
    for i in range(len(batches)):
        if i + 1 < len(batches):
            nb = batches[i + 1]
            slots.append(engine.submit(nb.cell, nb.gene, nb.value))
        outs.append(engine.collect(slots[i], shuffle_seed=42 + i, **kw))
    for i, (b, got) in enumerate(zip(batches, outs)):
        want = c_oracle.tokenize(b.cell, b.gene, b.value, "scgpt_binned", 512, perm=py_perm(42 + i, b.n_cells), n_bins=51, vocab_size=65000)
        compare(got, want, "scgpt_binned")


def test_slots_out_of_order_collect_and_exhaustion(engine):
    """Eight batches may be in flight; they can be collected in any order; a ninth submit is refused (SLAFB_ERR_BUSY) and
    leaves the pipeline intact; a released slot is reused."""
    batches = [synthetic.make_batch(120 + 10 * i, 20000, 300, seed=300 + i) for i in range(9)]
    kw = dict(mode="geneformer", max_genes=256)
    slots = [engine.submit(b.cell, b.gene, b.value, shuffle_seed=7 + i) for i, b in enumerate(batches[:8])]
    assert sorted(slots) == list(range(8))
    with pytest.raises(_native.SlafB200Error) as ei:
        engine.submit(batches[8].cell, batches[8].gene, batches[8].value)
    assert ei.value.code == _native.ERR_BUSY
    order = [5, 0, 7, 2, 1, 6, 3, 4]
    outs = {i: engine.collect(slots[i], shuffle_seed=7 + i, **kw) for i in order}
    again = engine.submit(batches[8].cell, batches[8].gene, batches[8].value)
    assert again in slots
    outs[8] = engine.collect(again, shuffle_seed=15, **kw)
    for i, b in enumerate(batches):
        want = c_oracle.tokenize(b.cell, b.gene, b.value, "geneformer", 256, perm=py_perm(7 + i, b.n_cells))
        compare(outs[i], want, "geneformer")


def test_output_capacity_retry_and_errors(engine):
    b = synthetic.make_batch(200, 5000, 100, seed=9)
    got = engine.collect(engine.submit(b.cell, b.gene, b.value), mode="geneformer", max_genes=64, capacity_cells=3)
    want = c_oracle.tokenize(b.cell, b.gene, b.value, "geneformer", 64)
    compare(got, want, "geneformer")
    with pytest.raises(ValueError, match="max_genes must be >= 1"):
        engine.tokenize(b.cell, b.gene, b.value, mode="geneformer", max_genes=0)
    with pytest.raises(TypeError):
        engine.tokenize(b.cell.astype(np.int64), b.gene, b.value, mode="geneformer", max_genes=8)
    with pytest.raises(_native.SlafB200Error) as ei:
        engine.tokenize(b.cell, b.gene, b.value, mode="geneformer_pct", max_genes=8, min_percentile=150.0)
    assert ei.value.code == _native.ERR_INVALID
    # empty frame -> zero rows, not an error (reference tests/test_aggregators.py:133-144)
    e = engine.tokenize(np.zeros(0, np.uint32), np.zeros(0, np.uint16), np.zeros(0, np.uint16), mode="geneformer", max_genes=8)
    assert e.n_cells == 0 and tuple(e.input_ids.shape) == (0, 8)


# ------------------------------------------------------------------------ facades
def _lists(offs, flat):
    return [flat[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


@pytest.mark.parametrize("case", KATS, ids=[c["cite"][:48] for c in KATS])
def test_window_facade_reference_kats(engine, case):
    cell = np.asarray(case["cell"], np.int32)
    gene = np.asarray(case["gene"], np.int32)
    value = np.asarray(case["value"], case["value_dtype"])
    w = engine.window(cell, gene, value, mode=case["mode"], max_items=case["max_items"], n_bins=case["n_bins"],
                      min_percentile=case.get("min_percentile"))
    offs = w.offsets.cpu().numpy()
    assert w.cell_ids.cpu().tolist() == case["expect_cells"]
    assert _lists(offs, w.genes.cpu().numpy()) == case["expect_genes"]
    if "expect_exprs" in case:
        assert _lists(offs, w.exprs.cpu().numpy()) == [[float(x) for x in r] for r in case["expect_exprs"]]
    if "expect_bin_range" in case:
        lo, hi = case["expect_bin_range"]
        assert all(lo <= x < hi for x in w.exprs.cpu().tolist())


@pytest.mark.parametrize("value_dtype", ["uint16", "float32"])
@pytest.mark.parametrize("mode,pct", MODES)
def test_window_facade_vs_oracle(engine, mode, pct, value_dtype):
    k = 48
    cell, gene, value = adversarial_batch(31, value_dtype, "int32", "int32", k=k)
    C = len(np.unique(cell))
    perm = py_perm(7, C)
    cids, offs, genes, exprs = c_oracle.window(cell, gene, value, mode, k, perm=perm, n_bins=51, min_percentile=pct)
    w = engine.window(cell, gene, value, mode=mode, max_items=k, n_bins=51, min_percentile=pct, shuffle_seed=7)
    assert_same(w.cell_ids.cpu().numpy(), cids, "cell ids")
    assert_same(w.offsets.cpu().numpy(), offs, "offsets")
    assert_same(w.genes.cpu().numpy(), genes, "genes")
    if mode.startswith("scgpt"):
        assert_same(w.exprs.cpu().numpy(), exprs, "exprs")


def test_float32_window_bins_follow_the_hosts_log1pf(engine):
    """Binned window on float32 storage (aggregators.py:96-105): polars calls the platform's log1pf, the checker here calls the same
    function, and the device evaluates whichever of its two log1pf versions the host's libm matches (fdlibm restatement for glibc
    <= 2.39, correctly rounded for >= 2.40).  2^20 window bins make a one-ulp difference in log1pf visible: with the other version
    some bins would differ (asserted), with the selected one none does."""
    rng = np.random.default_rng(5)
    n, C = 4000, 24
    value = np.unique(rng.uniform(0.05, 200.0, n * C).astype(np.float32))
    n = value.size // C
    value = value[: n * C].copy()
    rng.shuffle(value)
    cell = np.repeat(np.arange(C, dtype=np.int32) + 3, n)
    gene = np.concatenate([np.sort(rng.choice(60000, n, replace=False)) for _ in range(C)]).astype(np.int32)
    nb = 1 << 20
    cids, offs, genes, exprs = c_oracle.window(cell, gene, value, "scgpt_binned", n + 1, n_bins=nb)
    w = engine.window(cell, gene, value, mode="scgpt_binned", max_items=n + 1, n_bins=nb)
    assert_same(w.offsets.cpu().numpy(), offs, "offsets")
    assert_same(w.genes.cpu().numpy(), genes, "genes")
    assert_same(w.exprs.cpu().numpy(), exprs, "exprs")
    # the test has teeth: the bins computed with the OTHER log1pf differ somewhere
    L = _native.lib()
    kind = L.slafb_debug_log1pf(-1, None, None, 0)
    lv = {}
    for kd in (0, 1):
        y = np.empty_like(value)
        L.slafb_debug_log1pf(kd, value.ctypes.data, y.ctypes.data, value.size)
        lv[kd] = y.reshape(C, n)
    def bins(v):
        m = v.max(axis=1, keepdims=True)
        return np.clip(np.floor((v * np.float32(nb)) / m), 0, nb - 1)
    assert np.array_equal(bins(lv[kind]).ravel(), np.asarray(exprs, np.float32).ravel())
    assert np.count_nonzero(bins(lv[kind]) != bins(lv[1 - kind])) > 0


@pytest.mark.parametrize("tag", ["gf8", "gf64", "gf2048"])
def test_tokenizer_facade_reference_golden_geneformer(engine, tag):
    offs, genes = to_dev(GOLD[f"{tag}_offsets"]), to_dev(GOLD[f"{tag}_genes"])
    ids, mask, vals = engine.tokenize_lists(offs, genes, None, scgpt=False, exprs_are_int=True, max_genes=int(GOLD[f"{tag}_max_genes"]))
    assert vals is None
    assert_same(ids.cpu().numpy(), GOLD[f"{tag}_ids"], "ids")
    assert_same(mask.cpu().numpy(), GOLD[f"{tag}_mask"], "mask")


@pytest.mark.parametrize("tag,is_int", [("sgi4", True), ("sgi1200", True), ("sgf_bins", False), ("sgf_bins10", False),
                                        ("sgf_raw", False), ("sgf_raw5", False)])
def test_tokenizer_facade_reference_golden_scgpt(engine, tag, is_int):
    mg, vocab, nb = (int(x) for x in GOLD[f"{tag}_cfg"])
    offs, genes, exprs = to_dev(GOLD[f"{tag}_offsets"]), to_dev(GOLD[f"{tag}_genes"]), to_dev(GOLD[f"{tag}_exprs"])
    ids, mask, vals = engine.tokenize_lists(offs, genes, exprs, scgpt=True, exprs_are_int=is_int, max_genes=mg, n_bins=nb, vocab_size=vocab)
    assert_same(ids.cpu().numpy(), GOLD[f"{tag}_ids"], "ids")
    assert_same(mask.cpu().numpy(), GOLD[f"{tag}_mask"], "mask")
    assert_same(vals.cpu().numpy(), GOLD[f"{tag}_vals"], "vals")


def test_gene_stats_vs_oracle(engine):
    b = synthetic.make_batch(400, 3000, 300, seed=77)
    s, c = engine.gene_stats(to_dev(b.gene), to_dev(b.value), 3000)
    want_s = np.bincount(b.gene.astype(np.int64), weights=b.value.astype(np.float64), minlength=3000).astype(np.int64)
    want_c = np.bincount(b.gene.astype(np.int64), minlength=3000)
    assert_same(s.cpu().numpy(), want_s, "sum")
    assert_same(c.cpu().numpy(), want_c, "count")


# ------------------------------------------------------------------------ full-size properties
def test_large_batch_properties_and_full_compare(engine):
    """A 4M-row prefetch batch (reference default prefetch_batch_size): size-independent
    properties on every row plus a full oracle comparison (the C oracle does this in seconds)."""
    b = synthetic.make_batch(2000, 60000, 2000, seed=2)
    kw = dict(mode="scgpt_binned", max_genes=1200, n_bins=51, vocab_size=65000)
    got = engine.tokenize(to_dev(b.cell), to_dev(b.gene), to_dev(b.value), shuffle_seed=42, **kw)
    ids, vals, mask = got.input_ids, got.values, got.attention_mask
    assert (ids[:, 0] == 1).all()
    assert ((ids != 0) == mask).all()
    assert ((ids == 2).sum(1) == 1).all()                       # exactly one SEP per scGPT row
    n_tok = mask.sum(1)
    sep_pos = (ids == 2).float().argmax(1)
    assert (sep_pos == n_tok - 1).all()                          # SEP is the last non-PAD token
    assert (vals[:, 0] == 0).all() and (vals[~mask] == 0).all()
    inner = vals[mask]
    assert ((inner == 0) | (inner == 65000 + 50)).all()          # binned path collapses to {PAD, top bin}
    assert sorted(got.cell_ids.cpu().tolist()) == list(range(2000))   # every cell exactly once
    want = c_oracle.tokenize(b.cell, b.gene, b.value, "scgpt_binned", 1200, perm=py_perm(42, 2000), n_bins=51, vocab_size=65000)
    compare(got, want, "scgpt_binned")


def test_prepared_shuffle_matches_inline_shuffle(engine):
    """slafb_prepare_shuffle: the helper thread's permutation is the one collect() would build itself;
    a mismatching seed at collect time falls back to the inline build."""
    b = synthetic.make_batch(700, 20000, 300, seed=12)
    kw = dict(mode="geneformer", max_genes=256)
    want = c_oracle.tokenize(b.cell, b.gene, b.value, "geneformer", 256, perm=py_perm(77, b.n_cells))
    slots = [engine.submit(b.cell, b.gene, b.value, shuffle_seed=77) for _ in range(2)]
    compare(engine.collect(slots[0], shuffle_seed=77, **kw), want, "geneformer")
    want2 = c_oracle.tokenize(b.cell, b.gene, b.value, "geneformer", 256, perm=py_perm(78, b.n_cells))
    compare(engine.collect(slots[1], shuffle_seed=78, **kw), want2, "geneformer")       # prepared for 77, collected with 78
    s = engine.submit(b.cell, b.gene, b.value, shuffle_seed=5)
    compare(engine.collect(s, **kw), c_oracle.tokenize(b.cell, b.gene, b.value, "geneformer", 256), "geneformer")   # prepared, then no shuffle


@pytest.mark.parametrize("cfg", ["cfg2_scgpt_binned", "cfg3_geneformer"])
def test_full_size_prefetch_batch_checksums(cfg):
    """BASELINE.json's bench shape at full prefetch-batch size (32,768 cells, ~70 M rows): properties of EVERY output
    row against numpy reductions of the input (a checksum per cell, no per-cell Python), plus the full oracle on a
    random sample of cells.  Count data has far fewer distinct values per cell than max_items, so the reference's
    dense-rank filter keeps every row and each cell's tokens are its first min(n, limit) genes in row order."""
    from slaf_b200.engine import HotPathEngine

    C = 32768
    if cfg == "cfg2_scgpt_binned":
        b = synthetic.make_batch(C, 60000, 2000, seed=2)
        kw = dict(mode="scgpt_binned", max_genes=1200, n_bins=51, vocab_size=65000)
        L, limit = 1202, 1200
    else:
        b = synthetic.make_batch(C, 62000, 2000, seed=3)
        kw = dict(mode="geneformer", max_genes=2048)
        L, limit = 2048, 2047
    eng = HotPathEngine(0, max_rows=b.n_rows + 1024, max_cells=C + 1024)
    got = eng.tokenize(to_dev(b.cell), to_dev(b.gene), to_dev(b.value), shuffle_seed=42, **kw)
    ids, mask, cids = got.input_ids, got.attention_mask, got.cell_ids
    assert tuple(ids.shape) == (C, L) and got.n_cells == C
    order = np.asarray(py_perm(42, C))
    assert_same(cids.cpu().numpy(), order, "cell order = CPython shuffle of first-appearance order")
    nnz = np.diff(b.cell_start)
    n_emit = np.minimum(nnz, limit)
    # per-cell checksum of the gene tokens: sum over the first n_emit rows of (gene + 4), via one cumulative sum
    csum = np.concatenate(([0], np.cumsum(b.gene.astype(np.int64) + 4)))
    want_sum = csum[b.cell_start[:-1] + n_emit] - csum[b.cell_start[:-1]] + 1 + np.where(n_emit + 1 < L, 2, 0)
    assert_same(ids.sum(1).cpu().numpy(), want_sum[order], "row checksums")
    assert_same(mask.sum(1).cpu().numpy(), np.minimum(n_emit + 2, L)[order], "attention lengths")
    assert bool(((ids != 0) == mask).all()) and bool((ids[:, 0] == 1).all())
    assert bool(((ids == 2).sum(1) == torch.from_numpy((n_emit + 1 < L)[order]).cuda()).all())          # SEP present iff it fits
    if got.values is not None:
        vals = got.values
        assert bool((vals[:, 0] == 0).all()) and bool((vals[~mask] == 0).all())
        assert bool(((vals == 0) | (vals == 65000 + 50)).all())
        # bin >= 1  <=>  floor(log1p(v) * 51 / log1p(vmax)) >= 1, evaluated with numpy on every kept row
        vmax = np.maximum.reduceat(b.value, b.cell_start[:-1])
        lv = np.log1p(b.value.astype(np.float64))
        top = (np.floor(lv * 51 / np.repeat(np.log1p(vmax.astype(np.float64)), nnz)) >= 1)
        tsum = np.concatenate(([0], np.cumsum(top.astype(np.int64))))
        want_top = tsum[b.cell_start[:-1] + n_emit] - tsum[b.cell_start[:-1]]
        assert_same((vals != 0).sum(1).cpu().numpy(), want_top[order], "top-bin counts")
    # the full oracle on a random sample of cells
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(C, 384, replace=False))
    rows = np.concatenate([np.arange(b.cell_start[c], b.cell_start[c + 1]) for c in pick])
    okw = {k: v for k, v in kw.items() if k != "mode"}
    want = c_oracle.tokenize(b.cell[rows], b.gene[rows], b.value[rows], kw["mode"], **okw)
    pos = np.empty(C, np.int64)
    pos[order] = np.arange(C)
    sel = torch.from_numpy(pos[pick]).cuda()
    assert_same(ids[sel].cpu().numpy(), want[0], "sampled ids")
    assert_same(mask[sel].cpu().numpy(), want[1], "sampled mask")
    if got.values is not None:
        assert_same(got.values[sel].cpu().numpy(), want[2], "sampled values")
    eng.close()


@pytest.mark.parametrize("mode", ["geneformer", "scgpt_raw", "scgpt_binned"])
def test_float32_distinct_count_boundaries(engine, mode):
    """float32 storage: the fast kernel decides "at most k distinct values" with an exact hash set (trusted up to
    1536 entries); cells just below / at / above k and around the set's limit must all match the oracle."""
    rng = np.random.default_rng(9)
    cases = []
    for k, distincts in ((100, (1, 99, 100, 101, 150)), (1536, (1535, 1536, 1537)), (2048, (1530, 1536, 1537, 2047, 2048, 2049, 2600))):
        for d in distincts:
            n = max(d, k) + 37
            vals = np.concatenate([np.arange(1, d + 1), rng.integers(1, d + 1, n - d)]).astype(np.float64)
            rng.shuffle(vals)
            cases.append((k, (vals * 0.37 + 0.01).astype(np.float32)))
    for k in (100, 1536, 2048):
        cols = [(i, v) for i, (kk, v) in enumerate(cases) if kk == k]
        cell = np.concatenate([np.full(len(v), 1000 + i, np.uint32) for i, v in cols])
        gene = np.concatenate([np.sort(rng.choice(60000, len(v), replace=False)) for _, v in cols]).astype(np.uint16)
        value = np.concatenate([v for _, v in cols])
        want = c_oracle.tokenize(cell, gene, value, mode, k, n_bins=51, vocab_size=65000)
        got = engine.tokenize(to_dev(cell), to_dev(gene), to_dev(value), mode=mode, max_genes=k, n_bins=51, vocab_size=65000)
        compare(got, want, mode)


@pytest.mark.parametrize("mode", ["geneformer", "scgpt_binned"])
def test_float32_many_distinct_values_switch_to_the_cta_kernel(mode):
    """float32 storage, warp-per-cell kernel (K2f): its per-warp set holds 255 distinct values; cells with more of them (but at most
    max_items) are deferred to the general kernel, and a batch that defers more than one cell in eight moves the context to the
    CTA-per-cell kernel (set of 1,536) for good.  Both batches must match the oracle; the deferral count shows the switch."""
    from slaf_b200.engine import HotPathEngine

    rng = np.random.default_rng(21)
    k, C, n, d = 350, 256, 400, 300
    cell = np.repeat(np.arange(C, dtype=np.uint32) + 5, n)
    gene = np.concatenate([np.sort(rng.choice(60000, n, replace=False)) for _ in range(C)]).astype(np.uint16)
    value = np.concatenate([rng.permutation(np.concatenate([np.arange(1, d + 1), rng.integers(1, d + 1, n - d)])) for _ in range(C)])
    value = (value * 0.37 + 0.01).astype(np.float32)
    want = c_oracle.tokenize(cell, gene, value, mode, k, n_bins=51, vocab_size=65000)
    eng = HotPathEngine(0, max_rows=cell.size + 1024, max_cells=C + 16)
    try:
        cols = (to_dev(cell), to_dev(gene), to_dev(value))
        slow = []
        for _ in range(3):
            compare(eng.tokenize(*cols, mode=mode, max_genes=k, n_bins=51, vocab_size=65000), want, mode)
            torch.cuda.synchronize()
            slow.append(eng.timing()["slow_cells"])
        if not os.environ.get("SLAFB_F32_CTA"):
            assert slow == [C, 0, 0], slow          # (timing() harvests the finished batch and returns the counts since the last call)
    finally:
        eng.close()


@pytest.mark.parametrize("seed", range(64))
def test_fuzz_random_shapes_and_parameters(engine, seed):
    """Randomised parity: random cell sizes (including 1-row and multi-thousand-row cells), value distributions
    (ties, heavy tails, all-equal, all-distinct), dtypes, modes, k / max_genes / n_bins, shuffles and perm selections."""
    rng = np.random.default_rng(1000 + seed)
    value_dtype = rng.choice(["uint16", "float32"])
    gene_dtype, cell_dtype = [("uint16", "uint32"), ("int32", "int32")][int(rng.integers(2))]
    G = 60000 if gene_dtype == "uint16" else 90000
    n_cells = int(rng.integers(1, 60))
    cells = []
    for _ in range(n_cells):
        n = int(rng.choice([1, 2, int(rng.integers(3, 200)), int(rng.integers(200, 3000)), int(rng.integers(3000, 9000))],
                           p=[0.1, 0.1, 0.4, 0.3, 0.1]))
        kind = rng.integers(5)
        if kind == 0:
            v = rng.geometric(0.3, n)
        elif kind == 1:
            v = np.full(n, int(rng.integers(1, 100)))
        elif kind == 2:
            v = rng.permutation(n) + 1
        elif kind == 3:
            v = rng.integers(1, max(2, n // 3), n)
        else:
            v = np.r_[rng.geometric(0.4, max(n - 2, 0)), rng.integers(1000, 60000, min(n, 2))][:n]
        cells.append((np.sort(rng.choice(G, n, replace=False)), v))
    ids = rng.choice(10_000_000, n_cells, replace=False)
    cell = np.concatenate([np.full(len(g), ids[i]) for i, (g, _) in enumerate(cells)]).astype(cell_dtype)
    gene = np.concatenate([g for g, _ in cells]).astype(gene_dtype)
    raw = np.concatenate([v for _, v in cells])
    value = np.minimum(raw, 65535).astype(np.uint16) if value_dtype == "uint16" else (np.sqrt(raw.astype(np.float64)) * 0.731).astype(np.float32)
    mode = str(rng.choice(["geneformer", "geneformer_pct", "scgpt_raw", "scgpt_binned"]))
    pct = float(rng.choice([0.0, 12.5, 50.0, 99.0, 100.0])) if mode == "geneformer_pct" else None
    max_genes = int(rng.choice([1, 2, 3, 17, 64, 255, 1200, 2048, 4097]))
    max_items = int(rng.choice([max_genes, 1, 5, 100, 3000]))
    n_bins = int(rng.choice([1, 2, 10, 51, 300]))
    kw = dict(max_items=max_items, n_bins=n_bins, vocab_size=int(rng.choice([1000, 65000])), min_percentile=pct)
    choice = int(rng.integers(3))
    if choice == 0:
        perm, okw = None, {}
    elif choice == 1:
        s = int(rng.integers(0, 100000))
        perm, okw = py_perm(s, n_cells), dict(shuffle_seed=s)
    else:
        perm = rng.permutation(n_cells)[: int(rng.integers(0, n_cells + 1))]
        okw = dict(perm=perm)
    want = c_oracle.tokenize(cell, gene, value, mode, max_genes, perm=None if choice == 2 else perm, **kw)
    if choice == 2:
        want = tuple(None if w is None else w[perm] for w in want)
    cols = (cell, gene, value) if rng.integers(2) else (to_dev(cell), to_dev(gene), to_dev(value))
    got = engine.tokenize(*cols, mode=mode, max_genes=max_genes, **kw, **okw)
    compare(got, want, mode)


def test_many_tiny_cells_take_the_dense_head_path(engine):
    """Hundreds of thousands of 1-3 row cells: every CSR-build span holds far more segment heads than its shared-memory
    list (the replay path with a block scan per tile), the fast kernel sees tiny segments, the permutation is large."""
    rng = np.random.default_rng(77)
    C = 120_000
    lens = rng.integers(1, 4, C)
    cell = np.repeat(np.arange(C, dtype=np.uint32) * 3 + 5, lens)
    gene = rng.integers(0, 60000, cell.size).astype(np.uint16)
    value = rng.integers(1, 9, cell.size).astype(np.uint16)
    for mode in ("geneformer", "scgpt_binned"):
        want = c_oracle.tokenize(cell, gene, value, mode, 6, perm=py_perm(3, C), n_bins=10, vocab_size=70000)
        got = engine.tokenize(to_dev(cell), to_dev(gene), to_dev(value), mode=mode, max_genes=6, n_bins=10, vocab_size=70000, shuffle_seed=3)
        compare(got, want, mode)
    # float32 values and int32 ids through the same shapes
    want = c_oracle.tokenize(cell.astype(np.int32), gene.astype(np.int32), value.astype(np.float32), "scgpt_raw", 2, n_bins=10, vocab_size=70000)
    got = engine.tokenize(cell.astype(np.int32), gene.astype(np.int32), value.astype(np.float32), mode="scgpt_raw", max_genes=2, n_bins=10, vocab_size=70000)
    compare(got, want, "scgpt_raw")


# ------------------------------------------------------------------------ round 2: percentile mode in the fast kernel
@pytest.mark.parametrize("gene_dtype", ["uint16", "int32"])
@pytest.mark.parametrize("pct", [0.0, 10.0, 37.5, 50.0, 99.9, 100.0])
def test_percentile_fast_kernel_shapes(engine, pct, gene_dtype):
    """MODE_GF_PCT on uint16 values runs in tokenize_cells_kernel (presence bitmap in shared memory + in-place compaction):
    cells with n <= k (the percentile rule still filters them), value ranges wider than the 8,192-bit bitmap (deferred to the
    general kernel), cells longer than the stages (rows read from global memory), every value equal, one row, ties at the
    threshold, k below and above the number of distinct values."""
    rng = np.random.default_rng(int(pct * 10) + 5)
    G = 60000 if gene_dtype == "uint16" else 70000
    cells = []

    def add(vals):
        n = len(vals)
        cells.append((np.sort(rng.choice(G, size=n, replace=False)), np.asarray(vals)))

    add([7])                                                      # one row
    add(np.full(300, 3))                                          # one distinct value
    add(rng.geometric(0.35, 900))                                 # count-like, n <= k
    add(rng.geometric(0.2, 2500))                                 # count-like, n > k
    add(np.r_[rng.geometric(0.35, 1500), 9000])                   # range > 8192: general kernel
    add(np.r_[rng.geometric(0.35, 1500), 8191 + 1])               # range exactly 8192: bitmap
    add(rng.integers(1, 4000, 3000))                              # ~2100 distinct values: real top-k cut
    add(rng.integers(1, 50, 4090))                                # just below the stage capacity
    add(rng.integers(1, 50, 4097))                                # just above: tail rows from global memory
    add(rng.integers(1, 3000, 9000))                              # far above, many distinct
    add(np.repeat(np.arange(1, 41), 25))                          # 40 distinct x 25: percentile lands on a tie group
    for _ in range(12):
        n = int(rng.integers(1, 3000))
        add(np.minimum(rng.geometric(float(rng.uniform(0.05, 0.6)), n), 65535))
    order = rng.permutation(len(cells))
    ids = rng.choice(500000, size=len(cells), replace=False)
    cell = np.concatenate([np.full(len(cells[j][0]), ids[j]) for j in order]).astype(np.uint32)
    gene = np.concatenate([cells[j][0] for j in order]).astype(gene_dtype)
    value = np.concatenate([cells[j][1] for j in order]).astype(np.uint16)
    C = len(cells)
    # the same cells as float32 storage (a monotone, injective image of the counts: same ranks, same ties) run in the float32
    # warp-per-cell kernel, whose set holds 255 distinct values: the wide cells are deferred to the general kernel
    value_f32 = (value.astype(np.float64) * 0.37 + 0.01).astype(np.float32)
    for val in (value, value_f32):
        for k, max_genes in ((2048, 2048), (64, 64), (20, 300), (2048, 100)):
            want = c_oracle.tokenize(cell, gene, val, "geneformer_pct", max_genes, max_items=k, perm=py_perm(3, C), min_percentile=pct)
            got = engine.tokenize(cell, gene, val, mode="geneformer_pct", max_genes=max_genes, max_items=k, min_percentile=pct, shuffle_seed=3)
            compare(got, want, "geneformer_pct")
    t = engine.timing()
    assert t["tokenize_ms"] > 0.0                                  # the fast kernel ran


def test_percentile_synthetic_config_full_compare(engine):
    b = synthetic.make_batch(3000, 62000, 2000, seed=77)
    want = c_oracle.tokenize(b.cell, b.gene, b.value, "geneformer_pct", 2048, perm=py_perm(5, b.n_cells), min_percentile=10.0)
    got = engine.tokenize(b.cell, b.gene, b.value, mode="geneformer_pct", max_genes=2048, min_percentile=10.0, shuffle_seed=5)
    compare(got, want, "geneformer_pct")
    assert engine.timing()["slow_cells"] <= 0.05 * b.n_cells        # only the heavy-tail cells (range > 8192) take the general kernel


# ------------------------------------------------------------------------ round 2: pipe, host segment tables
def test_pipe_matches_submit_collect(engine):
    """slafb_pipe_*: one call per step, outputs into a ring; every completed batch equals the oracle, in push order."""
    from slaf_b200.engine import PreparedInput

    batches = [synthetic.make_batch(250 + 40 * i, 30000, 500, seed=400 + i) for i in range(7)]
    kw = dict(mode="scgpt_binned", max_genes=300, n_bins=51, vocab_size=65000)
    pipe = engine.pipe(ring=2, capacity_cells=600, depth=3, **kw)
    dev = [tuple(to_dev(a) for a in (b.cell, b.gene, b.value)) for b in batches]
    done = []
    for i, cols in enumerate(dev):
        ri, n, seq = pipe.push(PreparedInput([cols]), 90 + i)
        assert (ri < 0) == (i < 3)
        if ri >= 0:
            out = pipe.batch(ri, n)
            done.append((seq, [t.clone() if t is not None else None for t in (out.input_ids, out.attention_mask, out.values, out.cell_ids)]))
    while True:
        ri, n, seq = pipe.pop()
        if ri < 0:
            break
        out = pipe.batch(ri, n)
        done.append((seq, [t.clone() if t is not None else None for t in (out.input_ids, out.attention_mask, out.values, out.cell_ids)]))
    assert [s for s, _ in done] == list(range(7))
    for (seq, (ids, mask, vals, cids)), b in zip(done, batches):
        want = c_oracle.tokenize(b.cell, b.gene, b.value, "scgpt_binned", 300, perm=py_perm(90 + seq, b.n_cells), n_bins=51, vocab_size=65000)
        assert_same(ids.cpu().numpy(), want[0], "ids")
        assert_same(mask.cpu().numpy(), want[1], "mask")
        assert_same(vals.cpu().numpy(), want[2], "vals")
        assert_same(cids.cpu().numpy(), want[3], "cell ids")
    # a batch larger than the ring entries is dropped with an error and does not leak its slot
    big = synthetic.make_batch(700, 30000, 100, seed=1)
    with pytest.raises(_native.SlafB200Error) as ei:
        for _ in range(4):
            pipe.push(PreparedInput([(big.cell, big.gene, big.value)]), 1)
    assert ei.value.code == _native.ERR_CAPACITY
    pipe.close()
    got = engine.tokenize(batches[0].cell, batches[0].gene, batches[0].value, shuffle_seed=90, **kw)     # the context is still usable
    assert got.n_cells == batches[0].n_cells


def test_pipe_host_segments_route_and_csi_submit(engine):
    """The cell-start-index routes: slafb_submit_segments (table from the caller), slafb_submit_csi (table derived in the
    library from the index and the piece positions) and the pipe with a prepared table all give the COO route's tensors."""
    from slaf_b200.engine import PreparedInput

    b = synthetic.make_batch(900, 30000, 400, seed=21, first_cell_id=100)
    csi = np.concatenate((np.zeros(100, np.int64), b.cell_start)).astype(np.int64)       # cells 0..99 have no rows
    kw = dict(mode="geneformer", max_genes=512)
    ref = engine.tokenize(b.cell, b.gene, b.value, shuffle_seed=5, **kw)
    cuts = [0, 1234, 1235, 200000, b.n_rows]
    pieces = [(b.cell[lo:hi], b.gene[lo:hi], b.value[lo:hi]) for lo, hi in zip(cuts[:-1], cuts[1:])]
    slot, n = engine.submit_csi(pieces, cuts[:-1], csi, check=2, shuffle_seed=5)
    assert n == b.n_cells
    st, ids = engine.segments(slot)                                   # host-built table: served without waiting for the GPU
    assert np.array_equal(st, b.cell_start.astype(np.uint32)) and np.array_equal(ids, np.arange(100, 1000, dtype=np.uint32))
    got = engine.collect(slot, shuffle_seed=5, **kw)
    for x, y in ((got.input_ids, ref.input_ids), (got.attention_mask, ref.attention_mask), (got.cell_ids, ref.cell_ids)):
        assert torch.equal(x, y)
    assert engine.timing()["csr_ms"] >= 0.0
    pipe = engine.pipe(ring=2, capacity_cells=1000, depth=0, **kw)
    ri, n, _ = pipe.push(PreparedInput([(None, b.gene, b.value)], seg_start=b.cell_start.astype(np.uint32), seg_cell=np.arange(100, 1000, dtype=np.uint32)), 5)
    out = pipe.batch(ri, n)
    assert torch.equal(out.input_ids, ref.input_ids) and torch.equal(out.cell_ids, ref.cell_ids)
    pipe.close()
    # a stale index is refused before anything is submitted
    bad = csi.copy()
    bad[101:-1] += 1
    with pytest.raises(_native.SlafB200Error, match="cell_start_index does not describe"):
        engine.submit_csi(pieces, cuts[:-1], bad, check=1)
    # so is a caller-supplied table that is not monotone or does not cover the rows (it would send bulk copies out of bounds)
    st_bad = b.cell_start.astype(np.uint32).copy()
    st_bad[10], st_bad[11] = st_bad[11], st_bad[10]
    for table in (st_bad, b.cell_start.astype(np.uint32)[:-1]):
        with pytest.raises((_native.SlafB200Error, ValueError)):
            engine.submit_segments([(b.gene, b.value)], table, np.arange(100, 1000, dtype=np.uint32))
    got = engine.tokenize(b.cell, b.gene, b.value, shuffle_seed=5, **kw)
    assert torch.equal(got.input_ids, ref.input_ids)


# ------------------------------------------------------------------------ round 2: packed wire format (f4)
@pytest.mark.parametrize("gene_dtype", ["uint16", "int32"])
@pytest.mark.parametrize("mode,pct", MODES)
def test_packed_route_equals_the_unpacked_routes(engine, mode, pct, gene_dtype):
    """slafb_submit_packed: ~2 bytes per row cross the bus and the tokenizing kernel decodes the blocks in shared memory; the
    tensors are those of the COO route and of the oracle.  Adversarial cells (more rows than the stage holds, values and gene
    gaps beyond a byte, one-row cells, real top-k cuts) in one piece and cut into three."""
    from slaf_b200.engine import pack_rows

    k = 64
    cell, gene, value = adversarial_batch(23, "uint16", gene_dtype, "uint32", k=k)
    heads = np.flatnonzero(np.r_[True, cell[1:] != cell[:-1]])
    starts = np.r_[heads, len(cell)].astype(np.int64)
    ids = cell[heads]
    C = len(ids)
    kw = dict(mode=mode, max_genes=k, n_bins=51, vocab_size=65000, min_percentile=pct)
    want = c_oracle.tokenize(cell, gene, value, mode, k, perm=py_perm(9, C), n_bins=51, vocab_size=65000, min_percentile=pct)
    whole = pack_rows(gene, value, starts, ids)
    got = engine.collect(engine.submit_packed([whole]), shuffle_seed=9, **kw)
    compare(got, want, mode)
    cuts = [0, C // 3, C // 3 + 1, C]
    pieces = [pack_rows(gene, value, starts[a:b + 1], ids[a:b], pin=True) for a, b in zip(cuts[:-1], cuts[1:])]
    got = engine.collect(engine.submit_packed(pieces, shuffle_seed=9), shuffle_seed=9, **kw)
    compare(got, want, mode)
    for mg in (1, 7, 2048):                                               # widths around the stage sizes
        kw2 = dict(kw, max_genes=mg, max_items=k)
        want2 = c_oracle.tokenize(cell, gene, value, mode, mg, max_items=k, perm=py_perm(9, C), n_bins=51, vocab_size=65000, min_percentile=pct)
        compare(engine.collect(engine.submit_packed([whole]), shuffle_seed=9, **kw2), want2, mode)


def test_packed_route_synthetic_batch_and_pipe(engine):
    from slaf_b200.engine import PreparedInput, pack_rows

    b = synthetic.make_batch(3000, 60000, 2000, seed=91)
    ids = b.cell[b.cell_start[:-1]]
    pk = pack_rows(b.gene, b.value, b.cell_start, ids, pin=True)
    assert pk.nbytes < 2.1 * b.n_rows + 48 * b.n_cells
    engine.timing()                                                       # (counters since the previous test)
    for mode, kw in (("scgpt_binned", dict(max_genes=1200, n_bins=51, vocab_size=65000)), ("geneformer", dict(max_genes=2048)),
                     ("geneformer_pct", dict(max_genes=2048, min_percentile=10.0)), ("scgpt_raw", dict(max_genes=512, n_bins=10, vocab_size=50000))):
        want = c_oracle.tokenize(b.cell, b.gene, b.value, mode, kw["max_genes"], perm=py_perm(4, b.n_cells), n_bins=kw.get("n_bins", 10),
                                 vocab_size=kw.get("vocab_size", 50000), min_percentile=kw.get("min_percentile"))
        compare(engine.collect(engine.submit_packed([pk]), shuffle_seed=4, mode=mode, **kw), want, mode)
    t = engine.timing()
    assert t["csr_ms"] == 0.0                                             # no CSR build: the blocks are the cell boundaries
    kw = dict(mode="scgpt_binned", max_genes=1200, n_bins=51, vocab_size=65000)
    pipe = engine.pipe(ring=2, capacity_cells=3100, depth=1, **kw)
    inp = PreparedInput([pk])
    assert pipe.push(inp, 4)[0] < 0
    ri, n, seq = pipe.push(inp, 5)
    want = c_oracle.tokenize(b.cell, b.gene, b.value, "scgpt_binned", 1200, perm=py_perm(4, b.n_cells), n_bins=51, vocab_size=65000)
    compare(pipe.batch(ri, n), want, "scgpt_binned")
    pipe.close()
    with pytest.raises(_native.SlafB200Error):
        engine.collect_csr(engine.submit_packed([pk]))
