"""HotPathEngine: thin Python owner of one libslaf_b200 context.

PyTorch is plumbing here -- it allocates the output tensors and supplies the CUDA stream;
every computation of the hot path happens in the sm_100a kernels behind the C ABI
(include/slaf_b200.h).  One engine per (feeder thread, device).
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from . import _native as N

_NP_DT = {np.dtype(np.uint16): N.U16, np.dtype(np.int32): N.I32, np.dtype(np.uint32): N.U32, np.dtype(np.float32): N.F32}
_TORCH_DT = {torch.uint16: N.U16, torch.int32: N.I32, torch.uint32: N.U32, torch.float32: N.F32}
MODES = {"geneformer": N.GENEFORMER, "geneformer_pct": N.GENEFORMER_PCT, "scgpt_raw": N.SCGPT_RAW, "scgpt_binned": N.SCGPT_BINNED}


@dataclass
class TokenizedBatch:
    """Device-resident result of one prefetch batch (rows already in shuffled order)."""

    input_ids: torch.Tensor              # int64 [C, L]
    attention_mask: torch.Tensor         # bool  [C, L]
    values: torch.Tensor | None          # int64 [C, L] (scGPT) or None
    cell_ids: torch.Tensor               # int64 [C]
    n_cells: int


@dataclass
class CsrBatch:
    """Device-resident raw batch: CSR over (cells in shuffled order) x genes."""

    crow_indices: torch.Tensor           # int64 [C + 1]
    col_indices: torch.Tensor            # int32 [nnz]  gene_integer_id
    values: torch.Tensor                 # uint16 | float32 [nnz]
    cell_ids: torch.Tensor               # int64 [C]
    n_cells: int

    def to_sparse_csr(self, n_genes: int, dtype=torch.float32) -> torch.Tensor:
        return torch.sparse_csr_tensor(self.crow_indices, self.col_indices.to(torch.int64), self.values.to(dtype),
                                       size=(self.n_cells, n_genes))


@dataclass
class WindowResult:
    cell_ids: torch.Tensor               # int64 [C]
    offsets: torch.Tensor                # int64 [C+1]
    genes: torch.Tensor                  # int64 [K]
    exprs: torch.Tensor | None           # float64 [K]


class _Column:
    """pointer + dtype code + location of one COO column (numpy array or torch tensor)."""

    def __init__(self, x: Any, allowed: tuple[int, ...], name: str):
        if isinstance(x, torch.Tensor):
            if not x.is_contiguous():
                x = x.contiguous()
            if x.dtype not in _TORCH_DT:
                raise TypeError(f"{name}: unsupported dtype {x.dtype}")
            self.code = _TORCH_DT[x.dtype]
            self.ptr = x.data_ptr()
            self.n = x.numel()
            self.location = N.LOC_DEVICE if x.is_cuda else N.LOC_HOST
            self.device_index = x.device.index if x.is_cuda else None
        else:
            x = np.ascontiguousarray(x)
            if x.dtype not in _NP_DT:
                raise TypeError(f"{name}: unsupported dtype {x.dtype}")
            self.code = _NP_DT[x.dtype]
            self.ptr = x.ctypes.data
            self.n = x.size
            self.location = N.LOC_HOST
            self.device_index = None
        if self.code not in allowed:
            raise TypeError(f"{name}: dtype code {self.code} not allowed for this column")
        self.keep = x


def _coo(cell: Any, gene: Any, value: Any) -> tuple[N.Coo, tuple]:
    c = _Column(cell, (N.U32, N.I32), "cell_integer_id")
    g = _Column(gene, (N.U16, N.I32, N.U32), "gene_integer_id")
    v = _Column(value, (N.U16, N.F32), "value")
    if not (c.n == g.n == v.n):
        raise ValueError(f"column lengths differ: {c.n}, {g.n}, {v.n}")
    if not (c.location == g.location == v.location):
        raise ValueError("COO columns must all be on the host or all on the device")
    coo = N.Coo(c.ptr or None, g.ptr or None, v.ptr or None, c.n, c.code, g.code, v.code, c.location)
    return coo, (c, g, v)


class HotPathEngine:
    """Owns a slafb_ctx.  Not thread-safe (the C context is single-threaded by contract)."""

    def __init__(self, device: int | torch.device | str = 0, max_rows: int = 1 << 25, max_cells: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("slaf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise RuntimeError(f"slaf_b200 computes on CUDA devices only, got {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.max_rows = int(max_rows)
        self.max_cells = int(max_cells if max_cells is not None else max_rows)
        self._L = N.lib()
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()  # make sure the primary context exists
            ctx = ctypes.c_void_p()
            N.check(self._L.slafb_create(ctypes.byref(ctx), self.device.index, self.max_rows, self.max_cells))
        self._ctx = ctx
        self._cap_hint = 0
        self._pending: dict[int, tuple] = {}
        # host columns of a collected batch stay referenced until the slot is submitted again: collect() only ENQUEUES the
        # tokenization behind the batch's H2D copy, so the copy engine may still be reading them when collect() returns
        # (include/slaf_b200.h, conventions); the next submit into that slot waits for the slot's previous batch first
        self._parked: dict[int, Any] = {}
        self._seg_buf = None

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        if getattr(self, "_ctx", None):
            for ref in getattr(self, "_pipes", []):
                pipe = ref()
                if pipe is not None:
                    pipe.close()
            self._L.slafb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ helpers
    def _params(self, mode: str, max_genes: int, max_items: int | None, n_bins: int, vocab_size: int,
                min_percentile: float | None, shuffle_seed: int | None, perm, check_contiguous: bool = False,
                window_n_bins: int | None = None, order: str = "row", norm: torch.Tensor | None = None) -> tuple[N.Params, Any]:
        if mode not in MODES:
            raise ValueError(f"unknown mode {mode!r}")
        if max_genes < 1:
            raise ValueError(f"max_genes must be >= 1, got {max_genes}")
        p = N.Params()
        p.mode = MODES[mode]
        p.max_genes = int(max_genes)
        p.max_items = int(max_items if max_items is not None else max_genes)
        p.n_bins = int(n_bins)
        p.vocab_size = int(vocab_size)
        p.min_percentile = float(min_percentile) if min_percentile is not None else 0.0
        p.flags = N.FLAG_CHECK_CONTIGUOUS if check_contiguous else 0
        p.window_n_bins = int(window_n_bins) if window_n_bins else 0
        if order not in ("row", "rank"):
            raise ValueError(f"order must be 'row' (the reference's) or 'rank', got {order!r}")
        p.order = N.ORDER_RANK if order == "rank" else N.ORDER_ROW
        if norm is not None:
            if not (isinstance(norm, torch.Tensor) and norm.is_cuda and norm.dtype == torch.float32 and norm.is_contiguous()):
                raise TypeError("norm must be a contiguous float32 CUDA tensor [n_genes]")
            p.norm = norm.data_ptr()
            p.n_norm = norm.numel()
        keep = None
        if perm is not None:
            keep = np.ascontiguousarray(np.asarray(perm, np.int32))
            p.shuffle = N.SHUFFLE_PERM
            p.perm = keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
            p.n_perm = keep.size
        elif shuffle_seed is not None:
            p.shuffle = N.SHUFFLE_SEED
            p.seed = int(shuffle_seed)
        else:
            p.shuffle = N.SHUFFLE_NONE
        return p, keep

    def _alloc_out(self, cap: int, L: int, scgpt: bool):
        dev = self.device
        ids = torch.empty((cap, L), dtype=torch.int64, device=dev)
        mask = torch.empty((cap, L), dtype=torch.bool, device=dev)
        vals = torch.empty((cap, L), dtype=torch.int64, device=dev) if scgpt else None
        cids = torch.empty((cap,), dtype=torch.int64, device=dev)
        out = N.Out(ids.data_ptr(), mask.data_ptr(), vals.data_ptr() if scgpt else None, cids.data_ptr(), cap)
        return out, (ids, mask, vals, cids)

    # ------------------------------------------------------------------ the hot path
    def submit(self, cell: Any, gene: Any, value: Any, shuffle_seed: int | None = None) -> int:
        """Enqueue H2D (host columns) + CSR build on the side stream; returns the slot.  With
        ``shuffle_seed`` the library's helper thread prepares that shuffle while the GPU works."""
        coo, keep = _coo(cell, gene, value)
        slot = ctypes.c_int32(-1)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_submit(self._ctx, ctypes.byref(coo), self._stream(), ctypes.byref(slot)), self._ctx)
            if shuffle_seed is not None:
                N.check(self._L.slafb_prepare_shuffle(self._ctx, slot.value, int(shuffle_seed)), self._ctx)
        self._took(slot.value, keep, coo.n_rows)
        return slot.value

    def _took(self, slot: int, keep: Any, n_rows: int) -> None:
        self._parked.pop(slot, None)          # the library waited for the slot's previous batch before reusing it
        self._pending[slot] = (keep, n_rows)

    def submit_pieces(self, pieces: list[tuple[Any, Any, Any]]) -> int:
        """submit() for a batch given as consecutive row blocks [(cell, gene, value), ...] -- no host-side
        concatenation; with pinned / registered blocks every copy is an async DMA."""
        if not pieces:
            raise ValueError("no pieces")
        coos, keeps, total = [], [], 0
        for cell, gene, value in pieces:
            coo, keep = _coo(cell, gene, value)
            coos.append(coo)
            keeps.append(keep)
            total += coo.n_rows
        arr = (N.Coo * len(coos))(*coos)
        slot = ctypes.c_int32(-1)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_submit_pieces(self._ctx, arr, len(coos), self._stream(), ctypes.byref(slot)), self._ctx)
        self._took(slot.value, keeps[0] + tuple(keeps[1:]), total)
        return slot.value

    def submit_segments(self, pieces: list[tuple[Any, Any]], seg_start: np.ndarray, seg_cell: np.ndarray, cell_signed: bool = False,
                        shuffle_seed: int | None = None) -> int:
        """submit() when the cell boundaries are already known on the host (from SLAFArray._cell_start_index): `pieces` are
        consecutive (gene, value) row blocks, `seg_start` uint32 [C + 1] row offsets into their concatenation, `seg_cell`
        the cell ids.  The cell column never crosses PCIe and no CSR build runs."""
        if not pieces:
            raise ValueError("no pieces")
        coos, keeps, total = [], [], 0
        for gene, value in pieces:
            g = _Column(gene, (N.U16, N.I32, N.U32), "gene_integer_id")
            v = _Column(value, (N.U16, N.F32), "value")
            if g.n != v.n or g.location != v.location:
                raise ValueError("gene / value blocks must have equal length and location")
            coos.append(N.Coo(None, g.ptr or None, v.ptr or None, g.n, N.I32 if cell_signed else N.U32, g.code, v.code, g.location))
            keeps.append((g, v))
            total += g.n
        seg_start = np.ascontiguousarray(seg_start, np.uint32)
        seg_cell = np.ascontiguousarray(seg_cell).view(np.uint32) if np.asarray(seg_cell).dtype.itemsize == 4 else np.ascontiguousarray(seg_cell, np.uint32)
        C = seg_cell.size
        if seg_start.size != C + 1:
            raise ValueError("seg_start must have one more entry than seg_cell")
        arr = (N.Coo * len(coos))(*coos)
        slot = ctypes.c_int32(-1)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_submit_segments(self._ctx, arr, len(coos), seg_start.ctypes.data, seg_cell.ctypes.data, C, self._stream(),
                                                  ctypes.byref(slot)), self._ctx)
            if shuffle_seed is not None:
                N.check(self._L.slafb_prepare_shuffle(self._ctx, slot.value, int(shuffle_seed)), self._ctx)
        cell_col = _Column(np.zeros(1, np.int32 if cell_signed else np.uint32), (N.U32, N.I32), "cell_integer_id")
        self._took(slot.value, (cell_col, keeps[0][0], keeps[0][1]) + tuple(keeps[1:]) + (seg_start, seg_cell), total)
        return slot.value

    def submit_csi(self, pieces: list[tuple[Any, Any, Any]], row_lo: list[int], csi: np.ndarray, check: int = 1,
                   shuffle_seed: int | None = None) -> tuple[int, int]:
        """submit() when the dataset's ``_cell_start_index`` describes the rows: `pieces` are consecutive (cell | None, gene, value)
        HOST row blocks, `row_lo[i]` the table row of piece i's first row, `csi` the int64 index.  The library derives the segment
        table itself (C, O(cells)), probes it against the cell column (`check`: 0 none, 1 sampled, 2 every head) and ships gene +
        value columns plus 8 B per cell; the cell column never crosses PCIe and no CSR build runs.  Returns (slot, n_segments)."""
        if not pieces:
            raise ValueError("no pieces")
        if csi.dtype != np.int64 or not csi.flags.c_contiguous:
            raise TypeError("cell start index must be a contiguous int64 array")
        coos, keeps, signed = _csi_pieces(pieces)
        total = sum(c.n_rows for c in coos)
        arr = (N.Coo * len(coos))(*coos)
        lo = (ctypes.c_int64 * len(coos))(*[int(x) for x in row_lo])
        slot, n = ctypes.c_int32(-1), ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_submit_csi(self._ctx, arr, len(coos), lo, csi.ctypes.data, csi.size, int(check), self._stream(),
                                             ctypes.byref(slot), ctypes.byref(n)), self._ctx)
            if shuffle_seed is not None:
                N.check(self._L.slafb_prepare_shuffle(self._ctx, slot.value, int(shuffle_seed)), self._ctx)
        cell_col = _Column(np.zeros(1, np.int32 if signed else np.uint32), (N.U32, N.I32), "cell_integer_id")
        self._took(slot.value, (cell_col, keeps[0][1], keeps[0][2]) + tuple(keeps[1:]) + (csi,), total)
        return slot.value, int(n.value)

    def submit_packed(self, pieces: list["PackedRows"], shuffle_seed: int | None = None) -> int:
        """submit() for rows in the packed wire format (``pack_rows``): ~2 bytes per row cross PCIe, the tokenizing kernel decodes
        the blocks in its shared-memory stage.  Everything downstream (collect, the pipe) is unchanged; the token tensors are
        identical to the unpacked routes."""
        inp = PreparedInput(pieces)
        slot = ctypes.c_int32(-1)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_submit_packed(self._ctx, inp.arr, inp.n_pieces, inp._sc, inp.cell_dtype, self._stream(), ctypes.byref(slot)), self._ctx)
            if shuffle_seed is not None:
                N.check(self._L.slafb_prepare_shuffle(self._ctx, slot.value, int(shuffle_seed)), self._ctx)
        cell_col = _Column(np.zeros(1, np.int32 if pieces[0].cell_signed else np.uint32), (N.U32, N.I32), "cell_integer_id")
        self._took(slot.value, (cell_col, inp), inp.n_rows)
        return slot.value

    def wait_loaded(self, slot: int) -> None:
        """Block until the H2D copy of the batch last submitted into `slot` has finished (its host columns may be reused)."""
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_wait_loaded(self._ctx, int(slot)), self._ctx)

    def collect(self, slot: int, *, mode: str, max_genes: int, max_items: int | None = None, n_bins: int = 10,
                vocab_size: int = 50000, min_percentile: float | None = None, shuffle_seed: int | None = None,
                perm=None, capacity_cells: int | None = None, check_contiguous: bool = False,
                window_n_bins: int | None = None, order: str = "row", norm: torch.Tensor | None = None) -> TokenizedBatch:
        try:
            p, pkeep = self._params(mode, max_genes, max_items, n_bins, vocab_size, min_percentile, shuffle_seed, perm,
                                    check_contiguous, window_n_bins, order, norm)
        except Exception:
            self._abandon(slot)          # bad arguments must not leave the submitted batch occupying its slot
            raise
        keep, n_rows = self._pending.pop(slot)
        scgpt = p.mode >= N.SCGPT_RAW
        L = max_genes + 2 if scgpt else max_genes
        cap = capacity_cells or self._cap_hint or max(64, min(self.max_cells, n_rows // 256 + 64))
        cap = max(1, min(cap, self.max_cells))
        n = ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            while True:
                out, tensors = self._alloc_out(cap, L, scgpt)
                rc = self._L.slafb_collect(self._ctx, slot, ctypes.byref(p), ctypes.byref(out), ctypes.byref(n), self._stream())
                if rc == N.ERR_CAPACITY and cap < n.value <= self.max_cells:
                    cap = int(n.value)  # outputs were too small: the slot kept its CSR, retry once
                    continue
                N.check(rc, self._ctx)
                break
        C = int(n.value)
        self._cap_hint = max(self._cap_hint, min(self.max_cells, C + C // 8 + 16))
        ids, mask, vals, cids = tensors
        self._parked[slot] = keep            # see __init__: the H2D copy of these columns may still be in flight
        del pkeep
        return TokenizedBatch(ids[:C], mask[:C], vals[:C] if scgpt else None, cids[:C], C)

    def _abandon(self, slot: int) -> None:
        if slot in self._pending:
            self._pending.pop(slot)
            p = N.Params()
            p.mode, p.max_genes, p.max_items, p.n_bins, p.shuffle, p.n_perm = N.GENEFORMER, 1, 1, 1, N.SHUFFLE_PERM, 0
            out = N.Out(None, None, None, None, 0)
            n = ctypes.c_int64(0)
            with torch.cuda.device(self.device):
                self._L.slafb_collect(self._ctx, slot, ctypes.byref(p), ctypes.byref(out), ctypes.byref(n), self._stream())

    def release(self, slot: int) -> None:
        """Give a submitted slot back without tokenizing it (an empty selection emits no rows)."""
        if slot in self._pending:
            try:
                self.collect(slot, mode="geneformer", max_genes=1, perm=np.zeros(0, np.int32), capacity_cells=1)
            except N.SlafB200Error:
                pass                                            # the library frees the slot on every non-retryable error

    def segments(self, slot: int) -> tuple[np.ndarray, np.ndarray]:
        """Segment table of a submitted batch on the host: (seg_start uint32 [C+1], seg_cell [C]) with
        cells in order of first appearance.  Blocks until the slot's H2D copy + CSR build are done."""
        keep, n_rows = self._pending[slot]
        cap = max(1, min(self.max_cells, n_rows))
        if self._seg_buf is None or self._seg_buf[0].size < cap + 1:
            self._seg_buf = (np.empty(cap + 1, np.uint32), np.empty(cap, np.uint32))
        start, cellid = self._seg_buf
        n = ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_segments(self._ctx, slot, ctypes.byref(n), start.ctypes.data, cellid.ctypes.data, cap), self._ctx)
        C = int(n.value)
        ids = cellid[:C].copy()
        if keep[0].code == N.I32:
            ids = ids.view(np.int32)
        return start[: C + 1].copy(), ids

    def collect_csr(self, slot: int, *, shuffle_seed: int | None = None, perm=None) -> "CsrBatch":
        """Raw mode on the device: the submitted batch as CSR tensors, rows = cells in shuffled order
        (reference RawPrefetchBatch, slaf/ml/datasets.py:958-1008)."""
        try:
            p, pkeep = self._params("geneformer", 1, 1, 1, 0, None, shuffle_seed, perm)
        except Exception:
            self._abandon(slot)
            raise
        keep, n_rows = self._pending.pop(slot)
        n_out = len(pkeep) if pkeep is not None else None
        cap_cells = max(1, min(self.max_cells, n_rows) if n_out is None else n_out)
        dev = self.device
        crow = torch.zeros(cap_cells + 1, dtype=torch.int64, device=dev)
        col = torch.empty(max(n_rows, 1), dtype=torch.int32, device=dev)
        vdt = torch.uint16 if (len(keep) < 3 or getattr(keep[2], "code", N.U16) == N.U16) else torch.float32     # (packed batches: refused by the library below)
        val = torch.empty(max(n_rows, 1), dtype=vdt, device=dev)
        cids = torch.empty(cap_cells, dtype=torch.int64, device=dev)
        n, r = ctypes.c_int64(0), ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_collect_csr(self._ctx, slot, ctypes.byref(p), crow.data_ptr(), col.data_ptr(), val.data_ptr(),
                                              cids.data_ptr(), cap_cells, max(n_rows, 1), ctypes.byref(n), ctypes.byref(r), self._stream()), self._ctx)
        del keep, pkeep                      # (slafb_collect_csr synchronises the stream: the copy has finished)
        C, R = int(n.value), int(r.value)
        return CsrBatch(crow[: C + 1], col[:R], val[:R], cids[:C], C)

    def tokenize(self, cell: Any, gene: Any, value: Any, **kw) -> TokenizedBatch:
        """shuffle -> window -> tokenize -> cell ids for one batch of complete cells
        (reference: slaf/ml/datasets.py:1011-1098)."""
        self._params(kw["mode"], kw["max_genes"], kw.get("max_items"), kw.get("n_bins", 10), kw.get("vocab_size", 50000),
                     kw.get("min_percentile"), kw.get("shuffle_seed"), kw.get("perm"), kw.get("check_contiguous", False),
                     kw.get("window_n_bins"), kw.get("order", "row"), kw.get("norm"))      # raises before a slot is taken
        return self.collect(self.submit(cell, gene, value), **kw)

    # ------------------------------------------------------------------ facades
    def window(self, cell: Any, gene: Any, value: Any, *, mode: str, max_items: int, n_bins: int = 10,
               min_percentile: float | None = None, shuffle_seed: int | None = None, perm=None,
               check_contiguous: bool = False) -> WindowResult:
        """Window.apply as flat list columns (reference: slaf/ml/aggregators.py:26-50)."""
        coo, keep = _coo(cell, gene, value)
        p, pkeep = self._params(mode, max_items, max_items, n_bins, 0, min_percentile, shuffle_seed, perm, check_contiguous)
        R = coo.n_rows
        cap = max(1, min(self.max_cells, R))
        dev = self.device
        cids = torch.empty((cap,), dtype=torch.int64, device=dev)
        offs = torch.zeros((cap + 1,), dtype=torch.int64, device=dev)
        genes = torch.empty((max(R, 1),), dtype=torch.int64, device=dev)
        scgpt = p.mode >= N.SCGPT_RAW
        exprs = torch.empty((max(R, 1),), dtype=torch.float64, device=dev) if scgpt else None
        n = ctypes.c_int64(0)
        k = ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_window(self._ctx, ctypes.byref(coo), ctypes.byref(p), cids.data_ptr(), offs.data_ptr(),
                                         genes.data_ptr(), exprs.data_ptr() if scgpt else None, cap,
                                         ctypes.byref(n), ctypes.byref(k), self._stream()), self._ctx)
        del keep, pkeep
        C, K = int(n.value), int(k.value)
        return WindowResult(cids[:C], offs[: C + 1], genes[:K], exprs[:K] if scgpt else None)

    def tokenize_lists(self, offsets: torch.Tensor, genes: torch.Tensor, exprs: torch.Tensor | None, *, scgpt: bool,
                       exprs_are_int: bool, max_genes: int, n_bins: int = 10, vocab_size: int = 50000):
        """SLAFTokenizer.tokenize on flattened device lists (reference: slaf/ml/tokenizers.py:375-479, 635-722)."""
        B = offsets.numel() - 1
        L = max_genes + 2 if scgpt else max_genes
        if genes.numel() == 0:       # all lists empty: the kernel never dereferences, but NULL is refused by the ABI
            genes = torch.zeros(1, dtype=torch.int64, device=self.device)
        if scgpt and (exprs is None or exprs.numel() == 0):
            exprs = torch.zeros(1, dtype=torch.float64, device=self.device)
        out, (ids, mask, vals, _c) = self._alloc_out(max(B, 1), L, scgpt)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_tokenize_lists(self._ctx, int(scgpt), offsets.data_ptr(), genes.data_ptr(),
                                                 exprs.data_ptr() if (scgpt and exprs is not None) else None,
                                                 int(exprs_are_int), B, max_genes, n_bins, vocab_size,
                                                 ctypes.byref(out), self._stream()), self._ctx)
        return ids[:B], mask[:B], (vals[:B] if scgpt else None)

    def gene_stats(self, gene: torch.Tensor, value: torch.Tensor, n_genes: int, sum_out: torch.Tensor | None = None,
                   cnt_out: torch.Tensor | None = None):
        """Accumulate per-gene nonzero sum / count into uint64-as-int64 device tensors (extension)."""
        dev = self.device
        if sum_out is None:
            sum_out = torch.zeros(n_genes, dtype=torch.int64, device=dev)
        if cnt_out is None:
            cnt_out = torch.zeros(n_genes, dtype=torch.int64, device=dev)
        cell_dummy = torch.empty(gene.numel(), dtype=torch.int32, device=dev)
        coo, keep = _coo(cell_dummy, gene, value)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_gene_stats(self._ctx, ctypes.byref(coo), n_genes, sum_out.data_ptr(), cnt_out.data_ptr(),
                                             self._stream()), self._ctx)
        del keep
        return sum_out, cnt_out

    def gene_hist(self, gene: torch.Tensor, value: torch.Tensor, n_genes: int, n_bins: int = 1024, hist_out: torch.Tensor | None = None):
        """Accumulate the per-gene histogram of uint16 values into an int32 [n_genes, n_bins] device tensor (extension)."""
        if hist_out is None:
            hist_out = torch.zeros((n_genes, n_bins), dtype=torch.int32, device=self.device)
        cell_dummy = torch.empty(gene.numel(), dtype=torch.int32, device=self.device)
        coo, keep = _coo(cell_dummy, gene, value)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_gene_hist(self._ctx, ctypes.byref(coo), n_genes, n_bins, hist_out.data_ptr(), self._stream()), self._ctx)
        del keep
        return hist_out

    def gene_median(self, hist: torch.Tensor):
        """(median float32 [G], exact bool [G]) of the nonzero values per gene from a histogram table."""
        n_genes, n_bins = hist.shape
        med = torch.empty(n_genes, dtype=torch.float32, device=self.device)
        exact = torch.empty(n_genes, dtype=torch.bool, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_gene_median(self._ctx, hist.data_ptr(), n_genes, n_bins, med.data_ptr(), exact.data_ptr(),
                                              self._stream()), self._ctx)
        return med, exact

    def pipe(self, *, ring: int, capacity_cells: int, depth: int = 4, mode: str, max_genes: int, max_items: int | None = None,
             n_bins: int = 10, vocab_size: int = 50000, min_percentile: float | None = None, shuffle: bool = True,
             window_n_bins: int | None = None, order: str = "row", norm: torch.Tensor | None = None) -> "Pipe":
        """A steady pipeline with fixed parameters: one C call per step, outputs into a pre-allocated ring (slafb_pipe_*)."""
        p, _ = self._params(mode, max_genes, max_items, n_bins, vocab_size, min_percentile, 0 if shuffle else None, None, False,
                            window_n_bins, order, norm)
        import weakref

        pipe = Pipe(self, p, ring, capacity_cells, depth, norm)
        if not hasattr(self, "_pipes"):
            self._pipes = []
        self._pipes.append(weakref.ref(pipe))
        return pipe

    def set_timing_interval(self, every: int) -> None:
        """Take the per-kernel CUDA-event timings on every `every`-th batch only (0: never)."""
        N.check(self._L.slafb_set_option(self._ctx, N.OPT_TIMING_INTERVAL, int(every)), self._ctx)

    def set_front_on_caller(self, on: bool) -> None:
        """Device-resident batches: CSR build on the caller's stream (kernels never overlap)."""
        N.check(self._L.slafb_set_option(self._ctx, N.OPT_FRONT_ON_CALLER, int(bool(on))), self._ctx)

    def timing(self) -> dict[str, float]:
        t = N.Timing()
        with torch.cuda.device(self.device):
            N.check(self._L.slafb_get_timing(self._ctx, ctypes.byref(t)), self._ctx)
        return {f: getattr(t, f) for f, _ in N.Timing._fields_}


class PackedRows:
    """One batch (or a piece of one) in the packed wire format (include/slaf_b200.h, csrc/pack_rows.h): ``bytes`` uint8 (pinned when
    made by ``pack_rows(pin=True)``), ``blk_off`` uint32 [C + 1] (16-byte units), ``row_start`` uint32 [C + 1], ``seg_cell`` [C]
    cell ids, and the dtype the genes decode to."""

    def __init__(self, data: np.ndarray, blk_off: np.ndarray, row_start: np.ndarray, seg_cell: np.ndarray, gene_dtype: int, cell_signed: bool = False):
        self.bytes, self.blk_off, self.row_start = data, np.ascontiguousarray(blk_off, np.uint32), np.ascontiguousarray(row_start, np.uint32)
        sc = np.asarray(seg_cell)
        self.seg_cell = np.ascontiguousarray(sc).view(np.uint32) if sc.dtype.itemsize == 4 else np.ascontiguousarray(sc, np.uint32)
        self.gene_dtype, self.cell_signed = gene_dtype, cell_signed
        self.n_cells = int(self.seg_cell.size)
        self.n_rows = int(self.row_start[-1] - self.row_start[0]) if self.n_cells else 0
        if self.blk_off.size != self.n_cells + 1 or self.row_start.size != self.n_cells + 1:
            raise ValueError("blk_off / row_start need one more entry than there are cells")

    @property
    def nbytes(self) -> int:
        return int(self.bytes.nbytes)

    def struct(self) -> N.Packed:
        if isinstance(self.bytes, torch.Tensor):
            ptr, loc = self.bytes.data_ptr(), (N.LOC_DEVICE if self.bytes.is_cuda else N.LOC_HOST)
        else:
            ptr, loc = self.bytes.ctypes.data, N.LOC_HOST
        return N.Packed(ptr or None, self.nbytes, self.blk_off.ctypes.data, self.row_start.ctypes.data, self.n_cells, loc, self.gene_dtype)

    def unpack(self) -> tuple[np.ndarray, np.ndarray]:
        """(gene, value) columns back from the blocks (host; the library's reference decoder)."""
        data = self.bytes.cpu().numpy() if isinstance(self.bytes, torch.Tensor) else self.bytes
        gene = np.empty(self.n_rows, np.uint16 if self.gene_dtype == N.U16 else np.int32)
        value = np.empty(self.n_rows, np.uint16)
        rs = (self.row_start - self.row_start[0]).astype(np.uint32)
        N.check(N.lib().slafb_unpack_rows(data.ctypes.data, self.blk_off.ctypes.data, rs.ctypes.data, self.n_cells, gene.ctypes.data,
                                          self.gene_dtype, value.ctypes.data))
        return gene, value


def pack_rows(gene: np.ndarray, value: np.ndarray, row_start: np.ndarray, seg_cell: np.ndarray, pin: bool = False, cell_signed: bool = False) -> PackedRows:
    """Pack the cells whose rows are ``gene[row_start[c]:row_start[c+1]]`` / ``value[...]`` (host arrays; gene ids ascending inside a
    cell, uint16 values) into the packed wire format: ~2 bytes per row.  Single threaded C; pack disjoint cell ranges in parallel
    threads and submit them as pieces.  ``pin=True`` puts the blocks into page-locked memory (async DMA source)."""
    gene, value = np.ascontiguousarray(gene), np.ascontiguousarray(value)
    if value.dtype != np.uint16:
        raise TypeError("the packed format carries uint16 counts (float32 tables travel unpacked)")
    if gene.dtype not in _NP_DT or _NP_DT[gene.dtype] not in (N.U16, N.I32, N.U32):
        raise TypeError(f"gene_integer_id: unsupported dtype {gene.dtype}")
    rs = np.ascontiguousarray(row_start, np.int64)
    C = rs.size - 1
    if C < 0 or len(gene) != len(value):
        raise ValueError("bad row offsets / column lengths")
    lo = int(rs[0]) if C >= 0 and rs.size else 0
    rel = (rs - lo).astype(np.uint32)
    n_rows = int(rel[-1]) if rel.size else 0
    L = N.lib()
    cap = int(L.slafb_pack_bound(n_rows, C))
    while True:
        out = pinned_empty(cap, np.uint8) if pin else np.empty(cap + 16, np.uint8)
        if not pin:
            out = out[(-out.ctypes.data) % 16:][:cap]          # 16-byte aligned view
        blk = np.empty(C + 1, np.uint32)
        nb = ctypes.c_int64(0)
        g, v = gene[lo:lo + n_rows], value[lo:lo + n_rows]
        rc = L.slafb_pack_rows(g.ctypes.data, _NP_DT[gene.dtype], v.ctypes.data, rel.ctypes.data, C, out.ctypes.data, cap, blk.ctypes.data, ctypes.byref(nb))
        if rc == N.ERR_CAPACITY and cap < 20 * max(n_rows, 1) + 64 * max(C, 1) + 64:
            cap = 2 * cap + 64 * max(C, 1)                      # exception-heavy rows: retry with more room
            continue
        N.check(rc)
        break
    gdt = N.U16 if gene.dtype == np.uint16 else N.I32
    return PackedRows(out[: nb.value], blk, rel, seg_cell, gdt, cell_signed)


class PreparedInput:
    """The ctypes description of one batch's input (COO pieces, or gene / value pieces + a host segment table), built once for
    buffers that are submitted again and again (a feeder's staging ring, bench.py's resident batches)."""

    def __init__(self, pieces: list, seg_start: np.ndarray | None = None, seg_cell: np.ndarray | None = None,
                 cell_signed: bool = False):
        self.packed = bool(pieces) and isinstance(pieces[0], PackedRows)
        if self.packed:              # packed wire format: `pieces` are PackedRows, their cell ids are concatenated
            self.keep = list(pieces)
            self.arr = (N.Packed * len(pieces))(*[p.struct() for p in pieces])
            self.n_pieces = len(pieces)
            self.n_rows = sum(p.n_rows for p in pieces)
            self.seg_cell = np.ascontiguousarray(np.concatenate([p.seg_cell for p in pieces]))
            self._sc = self.seg_cell.ctypes.data
            self.cell_dtype = N.I32 if pieces[0].cell_signed else N.U32
            self.n_bytes = sum(p.nbytes for p in pieces)
            return
        coos, self.keep, self.n_rows = [], [], 0
        for cell, gene, value in pieces:
            if seg_start is None:
                coo, keep = _coo(cell, gene, value)
            else:
                g = _Column(gene, (N.U16, N.I32, N.U32), "gene_integer_id")
                v = _Column(value, (N.U16, N.F32), "value")
                if g.n != v.n or g.location != v.location:
                    raise ValueError("gene / value blocks must have equal length and location")
                coo, keep = N.Coo(None, g.ptr or None, v.ptr or None, g.n, N.I32 if cell_signed else N.U32, g.code, v.code, g.location), (g, v)
            coos.append(coo)
            self.keep.append(keep)
            self.n_rows += coo.n_rows
        self.arr = (N.Coo * len(coos))(*coos)
        self.n_pieces = len(coos)
        self.seg_start = self.seg_cell = None
        self.n_seg = 0
        if seg_start is not None:
            self.seg_start = np.ascontiguousarray(seg_start, np.uint32)
            self.seg_cell = np.ascontiguousarray(np.asarray(seg_cell)).view(np.uint32) if np.asarray(seg_cell).dtype.itemsize == 4 \
                else np.ascontiguousarray(seg_cell, np.uint32)
            self.n_seg = int(self.seg_cell.size)
            if self.seg_start.size != self.n_seg + 1:
                raise ValueError("seg_start must have one more entry than seg_cell")
        self._ss = self.seg_start.ctypes.data if self.seg_start is not None else None
        self._sc = self.seg_cell.ctypes.data if self.seg_cell is not None else None


class Pipe:
    """slafb_pipe_* behind torch-allocated ring tensors.  ``push`` returns ``(ring_index, n_cells, seq)`` of the batch the call
    completed (ring_index -1: none yet); ``batch`` gives that ring entry as tensor views.  Ring entry i is rewritten ``ring``
    completions later in stream order."""

    def __init__(self, eng: HotPathEngine, params: N.Params, ring: int, capacity_cells: int, depth: int, norm=None):
        self.eng, self._p, self._norm = eng, params, norm
        scgpt = params.mode >= N.SCGPT_RAW
        self.L = params.max_genes + 2 if scgpt else params.max_genes
        self.capacity_cells = int(capacity_cells)
        self.tensors, outs = [], []
        for _ in range(ring):
            out, t = eng._alloc_out(self.capacity_cells, self.L, scgpt)
            self.tensors.append(t)
            outs.append(out)
        self._ring = (N.Out * ring)(*outs)
        self._res = N.PipeResult()
        self._res_ref = ctypes.byref(self._res)
        self._pipe = ctypes.c_void_p()
        self._push, self._pop, self._push_packed = eng._L.slafb_pipe_push, eng._L.slafb_pipe_pop, eng._L.slafb_pipe_push_packed
        N.check(eng._L.slafb_pipe_create(eng._ctx, ctypes.byref(params), self._ring, ring, depth, ctypes.byref(self._pipe)), eng._ctx)
        self._stream = eng._stream()
        self._inputs: list[PreparedInput] = []        # referenced while their copies may be in flight (bounded by the slot count)

    def set_stream(self) -> None:
        """Re-read torch's current stream (the pipe launches on the stream that was current at creation / the last call of this)."""
        self._stream = self.eng._stream()

    def push(self, inp: PreparedInput, seed: int = 0) -> tuple[int, int, int]:
        if inp.packed:
            rc = self._push_packed(self._pipe, inp.arr, inp.n_pieces, inp._sc, inp.cell_dtype, seed, self._stream, self._res_ref)
        else:
            rc = self._push(self._pipe, inp.arr, inp.n_pieces, inp._ss, inp._sc, inp.n_seg, seed, self._stream, self._res_ref)
        if rc:
            N.check(rc, self.eng._ctx)
        self._inputs.append(inp)
        if len(self._inputs) > 16:
            del self._inputs[0]
        r = self._res
        return r.ring_index, r.n_cells, r.seq

    def pop(self) -> tuple[int, int, int]:
        rc = self._pop(self._pipe, self._stream, self._res_ref)
        if rc:
            N.check(rc, self.eng._ctx)
        r = self._res
        return r.ring_index, r.n_cells, r.seq

    def batch(self, ring_index: int, n_cells: int) -> TokenizedBatch:
        ids, mask, vals, cids = self.tensors[ring_index]
        return TokenizedBatch(ids[:n_cells], mask[:n_cells], vals[:n_cells] if vals is not None else None, cids[:n_cells], n_cells)

    def close(self) -> None:
        if self._pipe:
            self.eng._L.slafb_pipe_destroy(self._pipe)
            self._pipe = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def regroup_contiguous(cell: np.ndarray, gene: np.ndarray, value: np.ndarray):
    """Host control-plane helper: make every cell's rows one contiguous run, cells in order of
    first appearance, rows of a cell in row order -- what polars partition_by / group_by
    (maintain_order=True) do implicitly (reference slaf/ml/samplers.py:93).  The device kernels
    take contiguous cells; the host feeder calls this only when the library reports
    SLAFB_ERR_NONCONTIGUOUS (partial cells re-joined after a fragment boundary,
    reference slaf/ml/datasets.py:880-888)."""
    cell = np.asarray(cell)
    if cell.size == 0:
        return cell, np.asarray(gene), np.asarray(value)
    _, first, inv = np.unique(cell, return_index=True, return_inverse=True)
    order = np.argsort(first[inv], kind="stable")
    return cell[order], np.asarray(gene)[order], np.asarray(value)[order]


def _csi_pieces(pieces):
    coos, keeps, signed = [], [], False
    for cell, gene, value in pieces:
        g = _Column(gene, (N.U16, N.I32, N.U32), "gene_integer_id")
        v = _Column(value, (N.U16, N.F32), "value")
        c = _Column(cell, (N.U32, N.I32), "cell_integer_id") if cell is not None else None
        if g.n != v.n or g.location != v.location or (c is not None and (c.n != g.n or c.location != N.LOC_HOST)):
            raise ValueError("blocks must have equal length; the cell column (if given) must be on the host")
        signed = signed or (c is not None and c.code == N.I32)
        coos.append(N.Coo(c.ptr if c is not None and c.n else None, g.ptr or None, v.ptr or None, g.n,
                          c.code if c is not None else N.U32, g.code, v.code, g.location))
        keeps.append((c, g, v))
    if signed:
        for coo in coos:
            coo.cell_dtype = N.I32
    return coos, keeps, signed


def csi_segments(pieces: list[tuple[Any, Any, Any]], row_lo: list[int], csi: np.ndarray, check: int = 1) -> tuple[np.ndarray, np.ndarray]:
    """The segment table ``SLAFArray._cell_start_index`` implies for consecutive row blocks (the host half of ``submit_csi``;
    runs in the library's C code, no GPU involved): (seg_start uint32 [C + 1], seg_cell uint32 [C])."""
    if csi.dtype != np.int64 or not csi.flags.c_contiguous:
        raise TypeError("cell start index must be a contiguous int64 array")
    coos, keeps, _signed = _csi_pieces(pieces)
    rows = sum(c.n_rows for c in coos)
    cap = int(min(rows, max(csi.size - 1, 0) + len(coos))) + 1      # every piece can cut at most one more cell in two
    start, ids = np.empty(cap + 1, np.uint32), np.empty(cap, np.uint32)
    arr = (N.Coo * len(coos))(*coos)
    lo = (ctypes.c_int64 * len(coos))(*[int(x) for x in row_lo])
    n = ctypes.c_int64(0)
    N.check(N.lib().slafb_csi_segments(arr, len(coos), lo, csi.ctypes.data, csi.size, int(check), start.ctypes.data, ids.ctypes.data,
                                       cap, ctypes.byref(n)))
    del keeps
    C = int(n.value)
    return start[: C + 1].copy(), ids[:C].copy()


def shuffle_perm(seed: int, n: int) -> np.ndarray:
    """CPython random.seed(seed); random.shuffle(list(range(n))) via the library's host C code."""
    perm = np.empty(n, np.int32)
    N.check(N.lib().slafb_shuffle_perm(int(seed), int(n), perm.ctypes.data))
    return perm


def pinned_empty(n: int, dtype) -> np.ndarray:
    """numpy array over cudaHostAlloc'ed memory; freed when the last view of it is garbage collected."""
    import weakref

    dtype = np.dtype(dtype)
    ptr = ctypes.c_void_p()
    nbytes = max(1, n * dtype.itemsize)
    N.check(N.lib().slafb_host_alloc(ctypes.byref(ptr), nbytes))
    buf = (ctypes.c_char * nbytes).from_address(ptr.value)
    weakref.finalize(buf, _free_pinned, ptr.value)
    _PINNED_ALLOCS[ptr.value] = nbytes           # so that is_page_locked() recognises views of it
    return np.frombuffer(buf, dtype=dtype, count=n)


_PINNED_ALLOCS: dict[int, int] = {}     # base address -> bytes of live pinned_empty() allocations


def _free_pinned(address: int) -> None:
    _PINNED_ALLOCS.pop(address, None)
    try:
        N.lib().slafb_host_free(ctypes.c_void_p(address))
    except Exception:
        pass


# ---------------------------------------------------------------------------- registered (page-locked) user memory
_REGISTERED: dict[int, int] = {}     # base address -> bytes


def register_host_memory(arr: np.ndarray) -> bool:
    """Page-lock the memory of `arr` (cudaHostRegister) so slices of it are zero-copy DMA sources.
    Returns False when the driver refuses (e.g. read-only file mappings); the caller then keeps using
    the staged path."""
    if not isinstance(arr, np.ndarray) or arr.nbytes == 0 or not arr.flags.c_contiguous:
        return False
    base = arr.ctypes.data
    if is_page_locked(arr):
        return True
    rc = N.lib().slafb_host_register(ctypes.c_void_p(base), arr.nbytes)
    if rc != N.OK:
        return False
    _REGISTERED[base] = arr.nbytes
    return True


def unregister_host_memory(arr: np.ndarray) -> None:
    base = arr.ctypes.data
    if _REGISTERED.pop(base, None) is not None:
        N.lib().slafb_host_unregister(ctypes.c_void_p(base))


def register_host_range(address: int, nbytes: int) -> bool:
    """register_host_memory for a raw address range (e.g. a whole memory-mapped file: arrays that are views of it are then
    page-locked DMA sources).  False when the driver refuses."""
    if nbytes <= 0:
        return False
    if address in _REGISTERED and _REGISTERED[address] >= nbytes:
        return True
    if N.lib().slafb_host_register(ctypes.c_void_p(address), nbytes) != N.OK:
        return False
    _REGISTERED[address] = nbytes
    return True


def unregister_host_range(address: int) -> None:
    if _REGISTERED.pop(address, None) is not None:
        N.lib().slafb_host_unregister(ctypes.c_void_p(address))


def is_page_locked(arr: np.ndarray) -> bool:
    """True when `arr` lies inside memory registered through register_host_memory or allocated by pinned_empty."""
    a = arr.ctypes.data
    b = a + arr.nbytes
    for table in (_REGISTERED, _PINNED_ALLOCS):
        for base, size in table.items():
            if a >= base and b <= base + size:
                return True
    return False
