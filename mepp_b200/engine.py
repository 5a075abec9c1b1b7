"""Per-process profiling engine: owns the staged device buffers of one scored-FASTA
dataset and runs groups of (motif, orientation) tasks through the CUDA kernels.

Boundary: this replaces what, in the reference, is spread over
run_batch's save_dataset (mepp/batch.py:66-70), every run_single's load_dataset
(mepp/single.py:177) and dataset_and_motif_matrix_to_profile_data
(mepp/core.py:490-632): sequences are packed once into HBM, the score / permuted-score
operand Y is built once, and each task group is scan -> tcgen05 GEMM -> correlation
epilogue -> permutation statistics, with only L-length vectors returning to the host.

PyTorch is used for device memory, streams and (in parallel.py) torch.distributed only.
"""
import ctypes as C
import math
import os

import numpy as np

from . import _lib
from . import motif_layers
from . import stats_host

HIT_DTYPE = np.dtype([('task', np.int32), ('seq', np.int32), ('pos', np.int32), ('x0', np.float32)])


def _torch():
    import torch
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class ProfileBatch:
    """Results of a group of tasks: {column: (T, L)} positions, (T, N) density, thresholds, optional hits / r."""

    def __init__(self, task_ids, positions, density, scores, thr, hit_counts=None, hits=None, r=None, n_seq=0,
                 heatmaps=None, heatmap_spec=None):
        self.task_ids = task_ids
        self.positions = positions
        self.density = density
        self.scores = scores
        self.thr = thr
        self.hit_counts = hit_counts
        self.hits = hits
        self.r = r
        self.n_seq = n_seq
        # (T, H, W) float32 block maxima built on the device (profile(..., heatmap=(pixel_width, pixel_height, sigma)))
        # and the request they answer
        self.heatmaps = heatmaps
        self.heatmap_spec = heatmap_spec

    def __len__(self):
        return len(self.task_ids)

    def positions_df(self, i):
        import pandas as pd
        return pd.DataFrame({k: np.array(v[i]) for k, v in self.positions.items()})

    def ranks_df(self, i):
        """core.py:576-579,596-624: rank = arange (sort=False), score, density."""
        import pandas as pd
        return pd.DataFrame(dict(rank=np.arange(self.n_seq), score=self.scores,
                                 density=self.density[i].astype(np.float64)))

    def heatmap(self, i, L, pixel_width, pixel_height, sigma=1.0):
        """Clipped block-max of task i's score matrix on a pixel grid (plot.py:32-66, 106-145) from the match
        list, without materialising the (N, L) matrix."""
        from . import heatmap as _hm
        if self.heatmaps is not None and self.heatmap_spec == (pixel_width, pixel_height, sigma):
            return self.heatmaps[i]                     # reduced on the device (mepp_heatmap)
        if self.hits is None:
            raise RuntimeError("profile() was called without return_hits=True or heatmap=(width, height, sigma)")
        h = self.hits[self.hits['task'] == i]
        return _hm.compress_hits_to_pixels(h['seq'], h['pos'], h['x0'], (self.n_seq, L), pixel_width, pixel_height,
                                           sigma)

    def motif_score_matrix(self, i, L):
        """Dense (N, L) float32 unsmoothed score matrix of task i rebuilt from the match list
        (what core.py:568-573 returns as motif_score_matrix)."""
        if self.hits is None:
            raise RuntimeError("profile() was called without return_hits=True")
        h = self.hits[self.hits['task'] == i]
        X0 = np.zeros((self.n_seq, L), dtype=np.float32)
        X0[h['seq'], h['pos']] = h['x0']
        return X0


_COPY_POOLS = {}


def _copy_pool(n):
    from concurrent.futures import ThreadPoolExecutor
    if n not in _COPY_POOLS:
        _COPY_POOLS[n] = ThreadPoolExecutor(max_workers=n, thread_name_prefix="mepp-copy")
    return _COPY_POOLS[n]


SPARSE_DENSITY_MAX = 5e-3     # expected non-zero fraction of X up to which the sparse profile path is used (entry form)
SPARSE_MATCH_DENSITY_MAX = 1.5e-3   # expected matches per window up to which it is used (match-driven form)
SPARSE_ENTRY_DENSITY = 8e-3   # capacity of the entry list as a fraction of T * N * L
SPARSE_SMALL_ENTRIES = 4e6    # expected non-zeros per group below which the sparse path is used whatever the density
MIN_ENTRY_CAP = 1 << 20


def scan_units(tables):
    """Pair every two-filter ('+/-') task with a one-filter task that has the same filter 0, bias and length
    (the '+' orientation of the same motif: motif_layers.py:36-76 gives both the threshold of W): one pass of
    the scan kernel then serves both.  Returns int32 (U, 2) rows (task scanned, twin or -1), or None when no
    pair exists."""
    singles = {}
    for t, tb in enumerate(tables):
        if tb['n_filters'] == 1:
            key = (tb['m'], float(tb['bias']), tb['lut'][:, :, 0].tobytes())
            singles.setdefault(key, []).append(t)
    units, twinned = [], set()
    for t, tb in enumerate(tables):
        if tb['n_filters'] == 2:
            cand = singles.get((tb['m'], float(tb['bias']), tb['lut'][:, :, 0].tobytes()))
            if cand:
                u = cand.pop(0)
                twinned.add(u)
                units.append((t, u))
            else:
                units.append((t, -1))
    if not twinned:
        return None
    units.extend((t, -1) for t, tb in enumerate(tables) if tb['n_filters'] == 1 and t not in twinned)
    return np.array(units, dtype=np.int32)


class DeviceTables:
    """Scan tables of a task range [g0, g1) of a call whose thresholds were computed on the device."""

    def __init__(self, dev, host, g0, g1):
        self.dev, self.host, self.g0, self.g1 = dev, host, g0, g1
        self.mlen = host['mlen'][g0:g1]
        self.nfilt = host['nfilt'][g0:g1]
        self.pvalue = host['pvalue']

    def __len__(self):
        return self.g1 - self.g0

    def slice(self, g0, g1):
        return DeviceTables(self.dev, self.host, self.g0 + g0, self.g0 + g1)

    def device_slices(self):
        d, a, b = self.dev, self.g0, self.g1
        return d['lut'][a:b], d['bias'][a:b], d['mlen'][a:b], d['nfilt'][a:b], d['qtab'][a:b], d['qthr'][a:b]

    def bp_slice(self):
        """Records of the bit-parallel scan stage (mepp_scan_bitpar_tables) or None."""
        bp = self.dev.get('bp')
        return None if bp is None else bp[self.g0:self.g1]

    def units(self):
        """(task scanned, twin or -1) rows like scan_units, from the pairing made when the tables were prepared."""
        a, b = self.g0, self.g1
        twin = self.host['twin_of'][a:b]
        rows, twinned = [], set()
        for t in range(b - a):
            if self.nfilt[t] == 2:
                u = int(twin[t]) - a if a <= twin[t] < b else -1
                rows.append((t, u))
                if u >= 0:
                    twinned.add(u)
        if not twinned:
            return None
        rows.extend((t, -1) for t in range(b - a) if self.nfilt[t] == 1 and t not in twinned)
        return np.array(rows, dtype=np.int32)


class PendingProfile:
    """A profile_async() call whose last task group may still be on the GPU."""

    def __init__(self, eng, T_all, center, margin, ci, return_hits, return_r, on_group, n_groups):
        self.eng, self.T_all, self.center, self.margin, self.ci = eng, T_all, center, margin, ci
        self.return_hits, self.return_r, self.on_group, self.n_groups = return_hits, return_r, on_group, n_groups
        self.outs, self.paths, self.all_tables = [], [], []
        self.units = 0
        self.heatmap = None
        self.dtab = self.store = self.density_all = self.pos_view = self.batch = None

    def result(self):
        return self.eng._finish(self)


class Engine:
    def __init__(self, device=None, mem_budget_bytes=None, max_group=512):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("mepp_b200: no CUDA device visible; the engine has no CPU fallback")
        self.torch = torch
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        with torch.cuda.device(self.device):
            if not _lib.device_ok():
                raise RuntimeError("mepp_b200: device is not sm_100 (B200); kernels are built for sm_100a only")
        if mem_budget_bytes is None:
            # half of the device (90 GB on a B200): the staged dataset, the score operand (41 GB at cfg 5) and two
            # staging slots of per-group buffers have to fit
            with torch.cuda.device(self.device):
                mem_budget_bytes = torch.cuda.mem_get_info()[1] // 2
        self.mem_budget = int(mem_budget_bytes)
        self.max_group = int(max_group)
        self.staged = False
        self._bufs = {}
        self._pins = {}
        self._copy_stream = None
        self._aux_stream = None
        self.launches = 0   # kernels launched by this engine (bench.py reports it)
        self.timing = None  # set to {} to collect per-stage CUDA-event times (ms)
        self.host_timing = None   # set to {} to accumulate the host-side phases of profile() calls (ms)
        self.fuse_correlation = True   # False: separate GEMM (S) + correlation kernels (tests / A-B runs)
        self.share_scans = os.environ.get("MEPP_SHARE_SCANS", "1") != "0"   # '+' rides on the '+/-' pass of its motif
        self.profile_path = os.environ.get("MEPP_PROFILE_PATH", "auto")     # 'auto' | 'sparse' | 'tensor'
        self.last_paths = []
        self.fallbacks = []
        from .parallel import host_threads
        self.copy_threads = int(os.environ.get("MEPP_COPY_THREADS", str(max(1, min(4, host_threads())))))
        self.device_tables = os.environ.get("MEPP_DEVICE_TABLES", "1") != "0"   # threshold DP on the GPU
        # the scan of a group (integer-issue bound, no memory traffic) runs on the compute stream while the profile /
        # statistics / finalisation chain of the previous group (L2- and HBM-bound) runs on a high-priority stream
        self.overlap = os.environ.get("MEPP_OVERLAP", "1") != "0"
        self._post_stream = None
        self._post_done = {}
        self._carry = None          # the in-flight last group of a profile_async() call
        self._slot = 0              # staging slot of the next task group (round robin over ALL launches)
        # staging slots = task groups that may be on the GPU at once (their scan outputs and result copies are per
        # slot).  With four, every scan of a four-group call is enqueued before the host first waits for a result: in an
        # end-to-end step the scans then run under the upload of the permutation indices, which only the profile
        # chains need.
        self.n_slots = max(2, int(os.environ.get("MEPP_SLOTS", "4")))
        self._table_cache = {}      # task-list key -> DeviceTables (device_tables_for)
        # bit-parallel first stage of the scan for the units whose expected survivor rate allows it
        # (mepp_scan_bitpar); a unit takes it when its modelled cost is below MEPP_BP_COST_RATIO x the pair tables'
        self.bitpar = os.environ.get("MEPP_SCAN_BITPAR", "0") != "0"
        # lean scan for the sparse path (mepp_scan_lean): matches only, the blurred entries are built from buckets of
        # matches by a second kernel
        self.lean = os.environ.get("MEPP_SCAN_LEAN", "1") != "0"
        # tensor-core first stage of the lean scan (mepp_scan_tc: tcgen05 int8 products over the one-hot stream)
        self.tc_scan = os.environ.get("MEPP_SCAN_TC", "1") != "0"
        self.bp_cost_ratio = float(os.environ.get("MEPP_BP_COST_RATIO", "1.0"))
        # sparse profile from the match list (one row of Y per match; the blur commutes with the contraction:
        # mepp_profile_matches, per-position products S0 then their window sums) instead of the blurred entries (one
        # gathered row per non-zero, 2 margin + 1 times as many: mepp_profile_sparse).  cfg 3: 9.0 -> 4.9 ms per step.
        # MEPP_SPARSE_FORM=entries selects the entry form.
        self.sparse_matches = os.environ.get("MEPP_SPARSE_FORM", "matches") == "matches"

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _to_dev(self, a, dtype=None):
        torch = self.torch
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=True)

    def _buf(self, name, nbytes_or_shape, dtype, zero=False):
        """Grow-only named scratch buffer."""
        torch = self.torch
        shape = (int(nbytes_or_shape),) if np.isscalar(nbytes_or_shape) else tuple(int(s) for s in nbytes_or_shape)
        need = int(np.prod(shape))
        cur = self._bufs.get(name)
        if cur is None or cur.numel() < need or cur.dtype != dtype:
            if cur is not None:
                torch.cuda.synchronize(self.device)     # (rare: kernels of the side streams may still read the old one)
            cur = (torch.zeros if zero else torch.empty)(need, dtype=dtype, device=self.device)
            self._bufs[name] = cur
        return cur[:need].view(shape)

    # ------------------------------------------------------------------ staging
    def stage(self, ascii_u8, scores, perm_idx=None, use_gc=True):
        """Stage one dataset: ascii (N, L) uint8, scores (N,) float32 (rows already ordered as the
        reference's order_scored_fasta_df leaves them), perm_idx (P, N) int32 or None."""
        torch = self.torch
        self._drain()
        with torch.cuda.device(self.device):
            N, L = int(ascii_u8.shape[0]), int(ascii_u8.shape[1])
            if N < 4:
                raise ValueError("need at least 4 sequences")
            self.N, self.L = N, L
            self.n_pad = int(_lib.lib.mepp_pad32(N))
            W3 = int(_lib.lib.mepp_sym_words(L))
            st = self._stream()
            ascii_d = ascii_u8 if (isinstance(ascii_u8, torch.Tensor) and ascii_u8.is_cuda) else \
                self._to_dev(ascii_u8, torch.uint8).contiguous()
            self.sym_T = torch.empty((W3, self.n_pad), dtype=torch.int32, device=self.device)
            self.gc = torch.empty(self.n_pad, dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib.mepp_pack_dna(_ptr(ascii_d), N, L, _ptr(self.sym_T), _ptr(self.gc), st), "pack_dna")
            self.onehot = None
            if self.tc_scan and self.lean and L <= 65535:
                self.onehot = torch.empty(int(_lib.lib.mepp_onehot_words(N, L)), dtype=torch.int32, device=self.device)
                _lib.check(_lib.lib.mepp_pack_onehot(_ptr(ascii_d), N, L, _ptr(self.onehot), st), "pack_onehot")
                self.launches += 1
            self.planes_T = None
            if self.bitpar:
                WB = int(_lib.lib.mepp_plane_words(L))
                self.planes_T = torch.empty((WB, 4, self.n_pad), dtype=torch.int32, device=self.device)
                _lib.check(_lib.lib.mepp_pack_planes(_ptr(self.sym_T), N, L, _ptr(self.planes_T), st), "pack_planes")
                self.launches += 1
            self.launches += 2
            self.use_gc = bool(use_gc)
            if self.use_gc:
                # centre the covariate in float64 (Pearson r is shift invariant); plumbing, not a hot op
                z64 = self.gc[:N].double()
                zhat = torch.zeros(self.n_pad, dtype=torch.float32, device=self.device)
                zhat[:N] = (z64 - z64.mean()).float()
                self.zhat = zhat
                zh64 = zhat[:N].double()
                self._zsum_dev = torch.stack([zh64.sum(), (zh64 * zh64).sum()])
            else:
                self.zhat = None
                self._zsum_dev = None
            # scores: float32 like the reference's dataset (utils.py:188-192); centred + scaled on host in float64
            s32 = (scores.numpy() if isinstance(scores, torch.Tensor) else np.asarray(scores)).astype(np.float32)
            self.scores = s32
            y = s32.astype(np.float64)
            yc = y - y.mean()
            amax = float(np.abs(yc).max())
            scale = 1.0 if amax == 0.0 or not math.isfinite(amax) else 2.0 ** (13 - math.floor(math.log2(amax)))
            yhat = (yc * scale).astype(np.float32)
            yhat_d = self._to_dev(yhat)
            # the sparse path multiplies the float32 scores THEMSELVES (exact products, float64 sums: the raw-moment
            # cancellation of core.py:182-185 is harmless there); the centred / scaled copy is for the tensor path
            self._yraw_d = self._to_dev(s32)
            self._perm_src = None
            if perm_idx is not None and hasattr(perm_idx, 'rows'):
                # a permutations.PermutationSource: the index matrix is generated, uploaded and consumed in row chunks
                # when the score operand is built (_yrows); it never exists as a whole (cfg 5: 40 GB)
                if tuple(perm_idx.shape) != (int(perm_idx.shape[0]), N):
                    raise ValueError("perm_idx must have shape (P, N)")
                P, perm_d = int(perm_idx.shape[0]), None
                self._perm_src = perm_idx
                self._perm_ready = None
            elif isinstance(perm_idx, torch.Tensor) and perm_idx.is_cuda:
                P, perm_d = int(perm_idx.shape[0]), perm_idx          # already on the device (stage_broadcast)
                self._perm_ready = getattr(self, '_bcast_perm_ready', None)   # its broadcast may still be in flight
                self._bcast_perm_ready = None
            elif perm_idx is not None and int(perm_idx.shape[0]) > 0:
                P = int(perm_idx.shape[0])
                if tuple(perm_idx.shape) != (P, N):
                    raise ValueError("perm_idx must have shape (P, N)")
                # the permutation indices (the bulk of the staged bytes) travel on the side stream: only the Y
                # operands need them, and those are first used after the first scan of a call is in flight
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=self.device)
                src = perm_idx if isinstance(perm_idx, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(perm_idx))
                if src.dtype != torch.int32:
                    src = src.to(torch.int32)
                perm_d = self._buf('perm', (P, N), torch.int32)       # persistent: no cross-stream allocation
                free = torch.cuda.Event()
                free.record(torch.cuda.current_stream(self.device))   # earlier users of the buffer are done
                self._copy_stream.wait_event(free)
                with torch.cuda.stream(self._copy_stream):
                    perm_d.copy_(src, non_blocking=True)
                    self._perm_ready = torch.cuda.Event()
                    self._perm_ready.record(self._copy_stream)
            else:
                P, perm_d = 0, None
                self._perm_ready = None
            self.P, self.C = P, 1 + P
            self.BN, self.n_tiles_n = _lib.choose_bn(self.C)
            self.ldS = self.BN * self.n_tiles_n
            # Y operands are built at first use (_ensure_y)
            self._yhat_d, self._perm_d, self._y_rows = yhat_d, perm_d, None
            self._y_planes = self._col_syz = self._ysum = None
            self._col_syz_x = self._ysum_x = self._colconst_x = None      # statistics of the exact float32 operand
            self._zsum_host = None
            self._colconst_dev = None
            self.staged = True
        return self

    def stage_broadcast(self, ascii_u8=None, scores=None, perm_idx=None, use_gc=True, src=0):
        """Multi-GPU staging from ONE host copy: rank `src` passes the host arrays (the other ranks pass None), the
        device copies reach the other GPUs by NCCL broadcast over NVLink instead of one host upload per rank.  The
        permutation indices (the bulk of the bytes) are uploaded and broadcast on the side stream, behind the first
        scan (8-GPU end-to-end step on cfg 2: 38.3 -> 35.2 ms together with the leaner result gather)."""
        torch = self.torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self.stage(ascii_u8, scores, perm_idx, use_gc=use_gc)
        me = dist.get_rank() == src
        with torch.cuda.device(self.device):
            hdr = torch.zeros(3, dtype=torch.int64, device=self.device)
            if me:
                P = 0 if perm_idx is None else int(perm_idx.shape[0])
                hdr = torch.tensor([int(ascii_u8.shape[0]), int(ascii_u8.shape[1]), P], dtype=torch.int64,
                                   device=self.device)
            dist.broadcast(hdr, src)
            N, L, P = (int(v) for v in hdr.tolist())
            ascii_d = self._buf('ascii_bc', (N, L), torch.uint8)
            scores_d = self._buf('scores_bc', (N,), torch.float32)
            perm_d = self._buf('perm', (P, N), torch.int32) if P > 0 else None
            as_t = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dt)
            if me:
                ascii_d.copy_(as_t(ascii_u8, torch.uint8), non_blocking=True)
                scores_d.copy_(as_t(scores, torch.float32), non_blocking=True)
            dist.broadcast(ascii_d, src)
            dist.broadcast(scores_d, src)
            self._bcast_perm_ready = None
            if P > 0:
                # the permutation indices (the bulk of the bytes) travel on the side stream, upload and broadcast:
                # only the Y operands need them, and those are first used after the first scan is in flight
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=self.device)
                free = torch.cuda.Event()
                free.record(torch.cuda.current_stream(self.device))   # earlier users of the buffer are done
                self._copy_stream.wait_event(free)
                with torch.cuda.stream(self._copy_stream):
                    if me:
                        perm_d.copy_(as_t(perm_idx, torch.int32), non_blocking=True)
                    dist.broadcast(perm_d, src)
                    self._bcast_perm_ready = torch.cuda.Event()
                    self._bcast_perm_ready.record(self._copy_stream)
            return self.stage(ascii_d, scores_d.cpu().numpy(), perm_d, use_gc=use_gc)

    def stage_allgather(self, ascii_u8, scores, perm_idx=None, use_gc=True):
        """Multi-GPU staging when EVERY rank holds the host arrays (each rank parsed the same files, or generated the
        same synthetic dataset): rank k uploads only the k-th slice of the sequences and of the permutation indices
        through its own PCIe link, and the slices are all-gathered over NVLink / NVSwitch -- the host -> device bytes
        per rank shrink with the number of GPUs instead of staying at one full replica per rank (500 MB at cfg 3).
        The permutation indices travel on the side stream, behind the first scan."""
        torch = self.torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or hasattr(perm_idx, 'rows'):
            return self.stage(ascii_u8, scores, perm_idx, use_gc=use_gc)     # (a PermutationSource is staged per rank)
        world, rank = dist.get_world_size(), dist.get_rank()
        as_t = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dt)
        with torch.cuda.device(self.device):
            N, L = int(ascii_u8.shape[0]), int(ascii_u8.shape[1])
            P = 0 if perm_idx is None else int(perm_idx.shape[0])
            ra = -(-N // world)                                   # sequence rows per rank (the last slice may be short)
            ascii_all = self._buf('ascii_ag', (world * ra, L), torch.uint8)
            send_a = self._buf('ascii_ag_send', (ra, L), torch.uint8)
            a0, a1 = min(N, rank * ra), min(N, (rank + 1) * ra)
            if a1 > a0:
                send_a[:a1 - a0].copy_(as_t(ascii_u8, torch.uint8)[a0:a1], non_blocking=True)
            dist.all_gather_into_tensor(ascii_all, send_a)
            perm_d = None
            self._bcast_perm_ready = None
            if P > 0:
                rp = -(-P // world)
                perm_all = self._buf('perm', (world * rp, N), torch.int32)
                send_p = self._buf('perm_ag_send', (rp, N), torch.int32)
                p0, p1 = min(P, rank * rp), min(P, (rank + 1) * rp)
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=self.device)
                free = torch.cuda.Event()
                free.record(torch.cuda.current_stream(self.device))   # earlier users of the buffers are done
                self._copy_stream.wait_event(free)
                with torch.cuda.stream(self._copy_stream):
                    if p1 > p0:
                        send_p[:p1 - p0].copy_(as_t(perm_idx, torch.int32)[p0:p1], non_blocking=True)
                    dist.all_gather_into_tensor(perm_all, send_p)
                    self._bcast_perm_ready = torch.cuda.Event()
                    self._bcast_perm_ready.record(self._copy_stream)
                perm_d = perm_all[:P]
            s_host = scores.numpy() if isinstance(scores, torch.Tensor) else np.asarray(scores)
            return self.stage(ascii_all[:N], s_host, perm_d, use_gc=use_gc)

    def _ensure_y(self):
        """Tiled Y operand, column sums and sum of squares (mepp_build_y), once per staged dataset."""
        if self._ysum is None:
            torch = self.torch
            if self._perm_src is not None and self._perm_d is None:
                # the tensor path needs the whole index matrix on the device
                nbytes = 4 * self.P * self.N
                if nbytes > self.mem_budget // 4:
                    raise RuntimeError("the tensor profile path needs the whole permutation index matrix on the device "
                                       f"({nbytes >> 30} GB); use the sparse path or fewer permutations")
                self._perm_d = self._to_dev(self._perm_src.materialize())
            if self._perm_ready is not None:
                torch.cuda.current_stream(self.device).wait_event(self._perm_ready)
            self._y_planes = torch.empty(int(_lib.lib.mepp_y_planes_bytes(self.C, self.n_pad)), dtype=torch.uint8,
                                         device=self.device)
            self._col_syz = torch.empty(self.C, dtype=torch.float64, device=self.device)
            self._ysum = torch.empty(2, dtype=torch.float64, device=self.device)
            _lib.check(_lib.lib.mepp_build_y(_ptr(self._yhat_d), _ptr(self._perm_d), _ptr(self.zhat), self.N, self.P,
                                             _ptr(self._y_planes), _ptr(self._col_syz), _ptr(self._ysum), self._stream()),
                       "build_y")
            self.launches += 3 if self.use_gc else 2

    @property
    def y_planes(self):
        self._ensure_y()
        return self._y_planes

    @property
    def col_syz(self):
        self._ensure_y()
        return self._col_syz

    @property
    def ysum(self):
        self._ensure_y()
        return self._ysum

    PERM_STAGE_ROWS = 256      # permutation rows per staged chunk of a PermutationSource (1 GB at N = 1e6)

    def _build_y_chunked(self):
        """Score operand of the sparse path and its column statistics from a PermutationSource: the host generates
        chunk k + 1 (numpy shuffles on the host threads) while chunk k is uploaded from one of two pinned buffers and
        scattered into its columns of y_rows (mepp_build_y_cols)."""
        from concurrent.futures import ThreadPoolExecutor
        torch = self.torch
        src, N, P, C_ = self._perm_src, self.N, self.P, self.C
        ldY = int(_lib.lib.mepp_y_rows_ld(C_))
        self._y_rows = torch.empty((self.n_pad, ldY), dtype=torch.float32, device=self.device)
        self._col_syz_x = torch.empty(C_, dtype=torch.float64, device=self.device)
        self._ysum_x = torch.empty(2, dtype=torch.float64, device=self.device)
        st = self._stream()
        args = lambda perm, c0, nc: (_ptr(self._yraw_d), _ptr(perm), _ptr(self.zhat), N, c0, nc, C_, _ptr(self._y_rows),
                                     _ptr(self._col_syz_x), _ptr(self._ysum_x), st)
        _lib.check(_lib.lib.mepp_build_y_cols(*args(None, 0, 1)), "build_y_cols")           # column 0: the true scores
        self.launches += 3
        rows = self.PERM_STAGE_ROWS
        pins = [torch.empty((rows, N), dtype=torch.int32, pin_memory=True) for _ in range(2)]
        devs = [torch.empty((rows, N), dtype=torch.int32, device=self.device) for _ in range(2)]
        used = [None, None]
        bounds = [(p0, min(P, p0 + rows)) for p0 in range(0, P, rows)]
        with ThreadPoolExecutor(max_workers=1, thread_name_prefix="mepp-perm") as ex:
            gen = lambda k: src.rows(bounds[k][0], bounds[k][1], out=pins[k % 2].numpy()[:bounds[k][1] - bounds[k][0]])
            fut = ex.submit(gen, 0) if bounds else None
            for k, (p0, p1) in enumerate(bounds):
                fut.result()
                if k + 1 < len(bounds):
                    if used[(k + 1) % 2] is not None:
                        used[(k + 1) % 2].synchronize()      # the upload that read this pinned buffer is done
                    fut = ex.submit(gen, k + 1)
                devs[k % 2][:p1 - p0].copy_(pins[k % 2][:p1 - p0], non_blocking=True)
                used[k % 2] = torch.cuda.Event()
                used[k % 2].record(torch.cuda.current_stream(self.device))
                _lib.check(_lib.lib.mepp_build_y_cols(*args(devs[k % 2], 1 + p0, p1 - p0)), "build_y_cols")
                self.launches += 2
        torch.cuda.current_stream(self.device).synchronize()      # the staging buffers are freed on return

    def _ensure_y_exact(self):
        """Column statistics of the exact float32 score operand (sparse path: mepp_y_stats), once per staged dataset."""
        if self._ysum_x is None and self._perm_src is not None and self._perm_d is None:
            self._build_y_chunked()
        if self._ysum_x is None:
            torch = self.torch
            if self._perm_ready is not None:
                torch.cuda.current_stream(self.device).wait_event(self._perm_ready)
            self._col_syz_x = torch.empty(self.C, dtype=torch.float64, device=self.device)
            self._ysum_x = torch.empty(2, dtype=torch.float64, device=self.device)
            _lib.check(_lib.lib.mepp_y_stats(_ptr(self._yraw_d), _ptr(self._perm_d), _ptr(self.zhat), self.N, self.P,
                                             _ptr(self._col_syz_x), _ptr(self._ysum_x), self._stream()), "y_stats")
            self.launches += 2 if self.use_gc else 1

    def _yrows(self):
        if self._y_rows is None:
            self._ensure_y_exact()         # (a PermutationSource builds the operand together with its statistics)
        if self._y_rows is None:
            ldY = int(_lib.lib.mepp_y_rows_ld(self.C))
            self._y_rows = self.torch.empty((self.n_pad, ldY), dtype=self.torch.float32, device=self.device)
            _lib.check(_lib.lib.mepp_build_y_rows(_ptr(self._yraw_d), _ptr(self._perm_d), self.N, self.P,
                                                  _ptr(self._y_rows), self._stream()), "build_y_rows")
            self.launches += 1
        return self._y_rows

    def _bp_tables(self, lut_d, bias_d, mlen_d, nfilt_d, T):
        """Task records of the bit-parallel scan stage (letter sets, K, class) on the current stream, or None."""
        if not self.bitpar:
            return None
        bp = self.torch.empty((T, _lib.BP_WORDS), dtype=self.torch.int32, device=self.device)
        _lib.check(_lib.lib.mepp_scan_bitpar_tables(_ptr(lut_d), _ptr(bias_d), _ptr(mlen_d), _ptr(nfilt_d), T,
                                                    self.bp_cost_ratio, _ptr(bp), self._stream()), "scan_bitpar_tables")
        self.launches += 1
        return bp

    def device_tables_for(self, tasks, **table_kw):
        """Scan tables of a whole task list with the thresholds computed on the device: one host call for
        everything that does not need the threshold, one launch of the threshold DP over the distinct
        (motif, filter-0 orientation) units, one for the threshold-dependent table entries."""
        torch = self.torch
        # the tables depend on the task list alone (motif matrices, orientations, threshold parameters): a list that
        # comes again -- every step of a stream of datasets profiled against the same motifs -- reuses them
        key = self._table_key(tasks, table_kw)
        hit = self._table_cache.get(key)
        if hit is not None:
            return hit
        host = motif_layers.prepare_tables(tasks, **table_kw)
        T = len(tasks)
        rows = host['unit_rows']
        U = len(rows)
        # longest units first: their CTAs take the longest (the DP costs ~ m * score range)
        order = np.argsort(-host['dp_par'][rows, 7], kind='stable')
        rank_of = np.empty(U, dtype=np.int32)
        rank_of[order] = np.arange(U, dtype=np.int32)
        rows = rows[order]
        host = dict(host, unit_of_task=rank_of[host['unit_of_task']])
        R = host['dp_par'][rows, 7].astype(np.int64)
        off = np.zeros(U, dtype=np.int64)
        off[1:] = np.cumsum(2 * (R[:-1] + 1))
        n_scratch = int(off[-1] + 2 * (R[-1] + 1))
        # uploads and the two kernels run on an auxiliary stream, so that they overlap the staging copies of the
        # dataset that are still in flight on the compute stream; the compute stream waits for them before the scan
        compute = torch.cuda.current_stream(self.device)
        if self._aux_stream is None:
            self._aux_stream = torch.cuda.Stream(device=self.device)
        scratch = self._buf('dp_scratch', (n_scratch,), torch.float64)
        # (the scratch is free: the scans of the previous call waited for its DP, and that call has returned)
        with torch.cuda.stream(self._aux_stream):
            dev = dict(lut=self._to_dev(host['lut']), mlen=self._to_dev(host['mlen']), nfilt=self._to_dev(host['nfilt']),
                       qtab=self._to_dev(host['qtab'].view(np.int32)),
                       bias=torch.empty(T, dtype=torch.float32, device=self.device),
                       qthr=torch.empty((T, 2), dtype=torch.int32, device=self.device),
                       thr=torch.empty(T, dtype=torch.float64, device=self.device))
            thr_unit = torch.empty(U, dtype=torch.float64, device=self.device)
            st = self._stream()
            smat_d, umlen_d = self._to_dev(host['smat'][rows]), self._to_dev(host['mlen'][rows])
            par_d, off_d = self._to_dev(host['dp_par'][rows]), self._to_dev(off)
            uot_d, qpar_d = self._to_dev(host['unit_of_task']), self._to_dev(host['qpar'])
            _lib.check(_lib.lib.mepp_thresholds_dp(_ptr(smat_d), _ptr(umlen_d), _ptr(par_d), _ptr(off_d), _ptr(scratch),
                                                   _ptr(thr_unit), U, st), "thresholds_dp")
            _lib.check(_lib.lib.mepp_task_thresholds(_ptr(thr_unit), _ptr(uot_d), _ptr(qpar_d), _ptr(dev['mlen']),
                                                     _ptr(dev['nfilt']), T, _ptr(dev['bias']), _ptr(dev['qthr']),
                                                     _ptr(dev['thr']), st), "task_thresholds")
            bp = self._bp_tables(dev['lut'], dev['bias'], dev['mlen'], dev['nfilt'], T)
            if bp is not None:
                dev['bp'] = bp
            ready = torch.cuda.Event()
            ready.record(self._aux_stream)
        compute.wait_event(ready)
        for t_ in list(dev.values()) + [smat_d, umlen_d, par_d, off_d, uot_d, qpar_d, thr_unit]:
            t_.record_stream(compute)             # allocated on the auxiliary stream, read by the compute stream
        # (kept alive with the tables: the kernels above read them asynchronously)
        dev['_keep'] = (smat_d, umlen_d, par_d, off_d, uot_d, qpar_d, thr_unit)
        self.launches += 2
        out = DeviceTables(dev, host, 0, T)
        if len(self._table_cache) >= 4:
            self._table_cache.pop(next(iter(self._table_cache)))
        self._table_cache[key] = out
        return out

    @staticmethod
    def _table_key(tasks, table_kw):
        """Content key of a task list: the bytes of every distinct matrix object, the orientations, the parameters."""
        import hashlib
        h = hashlib.blake2b(digest_size=16)
        seen = {}
        for pwm, orient in tasks:
            d = seen.get(id(pwm))
            if d is None:
                a = np.ascontiguousarray(pwm, dtype=np.float64)
                d = seen[id(pwm)] = hashlib.blake2b(a.tobytes(), digest_size=16).digest() + bytes([a.shape[-1] & 255])
            h.update(d)
            h.update(str(orient).encode())
        bg = table_kw.get('bg')
        kw = tuple(sorted((k, (tuple(np.asarray(v, dtype=np.float64).ravel()) if k == 'bg' and v is not None else v))
                          for k, v in table_kw.items()))
        return h.digest(), len(tasks), kw

    def upload_tables(self, tables):
        """Host scan tables (motif_layers.scan_tables) -> DeviceTables resident in HBM, so that repeated profile()
        calls over the same task list move no table bytes."""
        torch = self.torch
        T = len(tables)
        host = dict(mlen=np.array([tb['m'] for tb in tables], dtype=np.int32),
                    nfilt=np.array([tb['n_filters'] for tb in tables], dtype=np.int32),
                    pvalue=max((tb.get('pvalue', 1e-4) for tb in tables), default=1e-4),
                    thr=np.array([tb['thr'] for tb in tables], dtype=np.float64))
        units = scan_units(tables)
        twin = np.full(T, -1, dtype=np.int32)
        if units is not None:
            for t, u in units:
                if u >= 0:
                    twin[t] = u
        host['twin_of'] = twin
        dev = dict(lut=self._to_dev(np.stack([tb['lut'] for tb in tables]).astype(np.float32)),
                   bias=self._to_dev(np.array([tb['bias'] for tb in tables], dtype=np.float32)),
                   mlen=self._to_dev(host['mlen']), nfilt=self._to_dev(host['nfilt']),
                   qtab=self._to_dev(np.stack([tb['qtab'] for tb in tables]).view(np.int32)),
                   qthr=self._to_dev(np.stack([tb['qthr'] for tb in tables]).view(np.int32)),
                   thr=self._to_dev(host['thr']))
        bp = self._bp_tables(dev['lut'], dev['bias'], dev['mlen'], dev['nfilt'], T)
        if bp is not None:
            dev['bp'] = bp
        return DeviceTables(dev, host, 0, T)

    def choose_path(self, tables, margin):
        """'sparse' (gather of Y rows per match, or per non-zero of X in the entry form) or 'tensor' (tcgen05 GEMM over
        the tiled operand).  The sparse path costs ~4 (1 + P) bytes of traffic per record, the GEMM the same whatever
        X holds: with the reference's p = 1e-4 threshold the sparse path wins by far; a dense X (large --motifpvalue
        with the flags honoured) goes to the tensor cores."""
        self._density_est = 0.0
        if self.profile_path in ('sparse', 'tensor'):
            return self.profile_path
        y_rows_bytes = 4 * self.n_pad * int(_lib.lib.mepp_y_rows_ld(self.C))
        if not self.fuse_correlation or y_rows_bytes > self.mem_budget // 2:
            return 'tensor'
        if isinstance(tables, DeviceTables):
            pv = tables.pvalue * int(tables.nfilt.max())
        else:
            pv = max((tb.get('pvalue', 1e-4) * tb['n_filters'] for tb in tables), default=1e-4)
        density = pv * (2 * int(margin) + 1)
        self._density_est = density
        # small problems always take the sparse path (it is exact in x and its cost is negligible there)
        small = density * len(tables) * self.N * self.L <= SPARSE_SMALL_ENTRIES
        if self.sparse_matches:
            # the match-driven form gathers one row of Y per MATCH, whatever the margin: measured on a cfg 2-sized
            # dataset (tools/gpu_dense_check.py) it is 3x faster than the GEMM at 1.6e-3 matches per window (p = 1e-3,
            # two strands, margin 5: 26 vs 84 us per task) and breaks even near 7e-3; the limit is set where the buckets
            # of the match publication (64 per task and 32 sequences) start to overflow as a rule
            return 'sparse' if pv <= SPARSE_MATCH_DENSITY_MAX or small else 'tensor'
        return 'sparse' if density <= SPARSE_DENSITY_MAX or small else 'tensor'

    def _colconst(self, sparse=False):
        """Per-dataset constants of the fused correlation epilogue (built once per stage()): from the statistics of the
        operand the path multiplies -- the exact float32 scores (sparse path) or their fp16 hi + lo split (tensor path)."""
        if sparse:
            if self._colconst_x is None:
                self._ensure_y_exact()
                cc = self.torch.empty(8 + 2 * self.C, dtype=self.torch.float64, device=self.device)
                _lib.check(_lib.lib.mepp_col_constants_dev(_ptr(self._col_syz_x), _ptr(self._ysum_x), _ptr(self._zsum_dev),
                                                           self.N, self.C, 1 if self.use_gc else 0, _ptr(cc),
                                                           self._stream()), "col_constants")
                self.launches += 1
                self._colconst_x = cc
            return self._colconst_x
        if self._colconst_dev is None:
            cc = self.torch.empty(8 + 2 * self.C, dtype=self.torch.float64, device=self.device)
            _lib.check(_lib.lib.mepp_col_constants_dev(_ptr(self.col_syz), _ptr(self.ysum), _ptr(self._zsum_dev), self.N,
                                                       self.C, 1 if self.use_gc else 0, _ptr(cc), self._stream()),
                       "col_constants")
            self.launches += 1
            self._colconst_dev = cc
        return self._colconst_dev

    def _zsum(self):
        if self._zsum_host is None:
            if self._zsum_dev is not None:
                self._zsum_host = self._zsum_dev.cpu().numpy().astype(np.float64)
            else:
                self._zsum_host = np.zeros(2, dtype=np.float64)
        return self._zsum_host

    # ------------------------------------------------------------------ sizing
    def group_size(self):
        """Tasks per group that fit the memory budget with the larger of the two profile paths (the tiled X operand)."""
        per_task = (self.L * self.n_pad * 4            # X planes (fp16 hi + lo; the sparse path needs ~1 % of this)
                    + self.L * self.ldS * 4            # r
                    + self.n_slots * (self.n_pad * 4 + self.L * 80))   # densities and finalised columns per staging slot
        g = max(1, min(self.max_group, self.mem_budget // max(per_task, 1)))
        return int(g)

    # ------------------------------------------------------------------ tasks
    def profile_async(self, tasks, margin=0, confidence_interval_pct=95, center=None, return_hits=False,
                      return_r=False, hits_cap=None, tables=None, honour_flags=False, pseudocount=1e-4, pvalue=1e-4,
                      bg=None, on_group=None, dp_precision=None, heatmap=None):
        """Enqueue (motif_matrix (4, m), orientation) tasks; returns a PendingProfile whose result() is the
        ProfileBatch.  The last task group of the call stays in flight: its device -> host copies and the host side of
        its results overlap whatever the caller enqueues next (the next profile_async call of a stream of task lists),
        and are completed by result() at the latest.  At most one call is left in flight.  on_group(t0, t1, positions,
        density), if given, is called as soon as the results of tasks t0 .. t1-1 are in the host tables (rows t0 .. t1-1
        of the {column: (T, L)} positions and of the (T, N) density are final): the caller can serialise them while
        the later groups are still on the GPU.  heatmap = (pixel_width, pixel_height, sigma): the clipped block maxima
        of every task's score matrix on that pixel grid (plot.py:32-66,106-145), reduced on the device from the match
        list (ProfileBatch.heatmaps / .heatmap(i, ...)); sigma None: no clip."""
        if not self.staged:
            raise RuntimeError("Engine.stage() must be called first")
        if not 0 <= int(margin) <= _lib.MAX_MARGIN:
            raise ValueError(f"margin must be in [0, {_lib.MAX_MARGIN}]")
        torch = self.torch
        T_all = len(tasks)
        L, N = self.L, self.N
        if center is None:
            center = L // 2
        if center < 0:
            center = 0
        if center > L:
            center = L - 1          # core.py:503-509
        tasks = list(tasks)
        for pwm, _o in tasks:
            if np.shape(pwm)[1] > L:
                raise ValueError("motif longer than the sequences")
        table_kw = dict(honour_flags=honour_flags, pseudocount=pseudocount, pvalue=pvalue, bg=bg)
        if dp_precision is not None:
            table_kw['dp_precision'] = float(dp_precision)       # grid of the MOODS threshold DP (motif_layers.DP_PRECISION)
        G = self.group_size()
        if T_all > 96:
            # several groups so that the host finalisation of one group overlaps the kernels of the next and the
            # exposed finalisation of the last group stays small (4 measured best on cfg 2) -- but no group below ~80
            # tasks: the tensor-core scan pays per N-tile of 16 scan units (32 tasks of a '+', '+/-' list), and 41
            # units fill three tiles where 21 units still need two (the per-rank share of a sharded task list)
            n_groups = int(os.environ.get("MEPP_GROUPS", "0")) or min(4, max(1, round(T_all / 81.0)))
            G = min(G, max(32, -(-T_all // n_groups)))
        G += G & 1 if G < T_all else 0      # even: the ('+', '+/-') pair of a motif stays in one group
        groups = [(g0, min(g0 + G, T_all)) for g0 in range(0, T_all, G)]
        call = PendingProfile(self, T_all, center, margin, confidence_interval_pct, return_hits, return_r, on_group,
                              len(groups))
        if heatmap is not None:
            call.heatmap = (heatmap[0], heatmap[1], heatmap[2] if len(heatmap) > 2 else 1.0)
        call.store = stats_host.alloc_positions(T_all, L, self.P)
        # matches per (task, sequence): <= L, so two bytes do whenever L <= 65535 (half the result bytes of a step)
        call.density_all = np.empty((T_all, N), dtype=np.uint16 if L <= 65535 else np.int32)
        call.pos_view = stats_host.positions_view(call.store, center) if on_group is not None else None
        import time as _time
        t_begin = _time.perf_counter()
        with torch.cuda.device(self.device):
            # the last group of the PREVIOUS call may still be in flight (profile_async): it is collected, like any
            # group, right after the next group has been enqueued, so the GPU never waits for the host side of a call
            inflight = [self._carry] if self._carry is not None else []
            self._carry = None
            # host threshold DP: every group's tables are submitted to the host thread pool now and collected when
            # the group is launched, so the DP of the later groups runs while the earlier ones are on the GPU
            dtab = futures = None
            if isinstance(tables, DeviceTables):
                dtab, tables = tables, None
            elif tables is None and self.device_tables:
                dtab = self.device_tables_for(tasks, **table_kw)
            elif tables is None:
                futures = [motif_layers.scan_tables_async(tasks[g0:g1], **table_kw) for g0, g1 in groups]
            call.dtab = dtab
            for gi, (g0, g1) in enumerate(groups):
                gt = dtab.slice(g0, g1) if dtab is not None else tables[g0:g1] if tables is not None else futures[gi]()
                if dtab is None:
                    call.all_tables.extend(gt)
                # round-robin slots: the group that used this slot n_slots launches ago has to be collected first
                while len(inflight) > self.n_slots - 1:
                    self._collect(inflight.pop(0))
                slot, self._slot = self._slot, (self._slot + 1) % self.n_slots
                nxt = self._launch_group(gt, margin, return_hits, return_r, hits_cap, confidence_interval_pct, slot,
                                         heatmap=call.heatmap)
                nxt['t0'] = g0
                nxt['call'] = call
                call.units += self.last_units
                call.paths.append(nxt['path'])
                inflight.append(nxt)
                # results that are already on the host are handed over at once (on_group: the file writers)
                while len(inflight) > 1 and inflight[0]['done'].query():
                    self._collect(inflight.pop(0))
                if return_hits or return_r:
                    # the match list / r buffers are reused by the next group: no overlap in this mode
                    while inflight:
                        self._collect(inflight.pop(0))
            while len(inflight) > 1:
                self._collect(inflight.pop(0))
            self._carry = inflight[0] if inflight else None
        if self.host_timing is not None:
            self.host_timing['enqueue_ms'] = self.host_timing.get('enqueue_ms', 0.0) + 1e3 * (_time.perf_counter() - t_begin)
        return call

    def profile(self, tasks, **kw):
        """Run (motif_matrix (4, m), orientation) tasks.  Returns a ProfileBatch (= profile_async(...).result())."""
        return self.profile_async(tasks, **kw).result()

    def _collect(self, pend):
        """Collect one in-flight group into the tables of the call it belongs to."""
        call = pend['call']
        o = self._collect_group(pend, call.margin, call.ci, call.store, call.density_all)
        call.outs.append(o)
        if call.on_group is not None:
            call.on_group(o['t0'], o['t0'] + o['T'], call.pos_view, call.density_all)

    def _drain(self):
        """Collect whatever is still in flight (before the staged buffers change)."""
        if self._carry is not None:
            pend, self._carry = self._carry, None
            with self.torch.cuda.device(self.device):
                self._collect(pend)

    def _finish(self, call):
        if call.batch is not None:
            return call.batch
        import time as _time
        if self._carry is not None and self._carry['call'] is call:
            t0 = _time.perf_counter()
            pend, self._carry = self._carry, None
            with self.torch.cuda.device(self.device):
                self._collect(pend)
            if self.host_timing is not None:
                # the exposed tail of a call: wait for the last group's copies, move its block into the result tables
                ht = self.host_timing
                ht['calls'] = ht.get('calls', 0) + 1
                ht['last_wait_ms'] = ht.get('last_wait_ms', 0.0) + 1e3 * (self._last_wait_end - t0)
                ht['last_copy_ms'] = ht.get('last_copy_ms', 0.0) + 1e3 * (_time.perf_counter() - self._last_wait_end)
        outs = call.outs
        if sum(o['T'] for o in outs) != call.T_all:
            raise RuntimeError("profile: a task group was never collected")
        cat = lambda k: np.concatenate([o[k] for o in outs], axis=0)
        # what the last call moved, for the roofline accounting of bench.py
        self.last_groups = call.n_groups
        self.last_paths = list(call.paths)
        self.last_stats = dict(hits=int(sum(int(o['hit_count'][0]) for o in outs)),
                               entries=int(sum(o['entries'] for o in outs)), units=int(call.units),
                               paths=list(call.paths))
        positions = stats_host.positions_view(call.store, call.center)
        hits = None
        if call.return_hits:
            parts = []
            off = 0
            for o in outs:
                h = o['hits'].copy()
                h['task'] += off
                off += o['T']
                parts.append(h)
            hits = np.concatenate(parts) if parts else np.zeros(0, dtype=HIT_DTYPE)
            order = np.lexsort((hits['pos'], hits['seq'], hits['task']))
            hits = hits[order]
        r = np.concatenate([o['r'] for o in outs], axis=0) if call.return_r else None
        dtab = call.dtab
        if dtab is not None:
            thr_all = dtab.host.get('_thr_np')
            if thr_all is None:
                thr_all = dtab.host['_thr_np'] = dtab.dev['thr'].cpu().numpy()     # (once per table set)
            thr = thr_all[dtab.g0:dtab.g1]
        else:
            thr = np.array([tb['thr'] for tb in call.all_tables])
        heat = np.concatenate([o['heatmap'] for o in outs], axis=0) if call.heatmap is not None else None
        call.batch = ProfileBatch(list(range(call.T_all)), positions, call.density_all, self.scores, thr,
                                  hit_counts=cat('hit_count'), hits=hits, r=r, n_seq=self.N, heatmaps=heat,
                                  heatmap_spec=call.heatmap)
        call.store = call.pos_view = call.on_group = None
        return call.batch

    def _reset_x_planes(self):
        """Restore the all-zero state of the A operand the last scan wrote (match-list driven, see
        mepp_x_planes_reset).  Deferred to the next use so that the operand can be inspected after profile()."""
        p = getattr(self, '_x_pending', None)
        if p is None:
            return
        self._x_pending = None
        _lib.check(_lib.lib.mepp_x_planes_reset(_ptr(p['x']), _ptr(p['hits']), p['cap'], _ptr(p['hit_count']),
                                                p['N'], p['L'], p['T'], p['margin'], self._stream()), "x_planes_reset")
        self.launches += 1

    def _mark(self, prev, name):
        """CUDA-event stage timer on the launching stream (only when self.timing is a dict)."""
        if self.timing is None:
            return None
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.torch.cuda.current_stream(self.device))
        if prev is not None:
            self.timing.setdefault('_pairs', []).append((name, prev, e))
        return e

    def collect_timing(self):
        """Sum the recorded stage times (ms) per stage name; call after a synchronize."""
        out = {}
        for name, a, b in (self.timing or {}).get('_pairs', []):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    def _pinned(self, name, slot, shape, dtype):
        """Grow-only pinned host buffer (two slots: one group in flight, one being finalised)."""
        torch = self.torch
        key = (name, slot)
        need = int(np.prod(shape))
        cur = self._pins.get(key)
        if cur is None or cur.numel() < need or cur.dtype != dtype:
            cur = torch.empty(max(need, 1), dtype=dtype, pin_memory=True)
            self._pins[key] = cur
        return cur[:need].view(tuple(int(x) for x in shape))

    def _launch_group(self, tables, margin, return_hits, return_r, hits_cap, ci, slot, path=None, lean=None,
                      matches=None, heatmap=None):
        """Enqueue every kernel of one task group plus the D2H copies of its results (no host sync)."""
        torch = self.torch
        T = len(tables)
        use_matches = self.sparse_matches if matches is None else bool(matches)
        path = path or self.choose_path(tables, margin)
        use_lean = self.lean if lean is None else bool(lean)
        L, N, C_, P = self.L, self.N, self.C, self.P
        st = self._stream()
        self._reset_x_planes()     # before any scratch buffer of the previous group is reused
        m_rows = T * L
        if isinstance(tables, DeviceTables):
            # tables prepared for the whole call and thresholded on the device (mepp_thresholds_dp): slices
            lut_d, bias_d, mlen_d, nfilt_d, qtab_d, qthr_d = tables.device_slices()
            mlen = tables.mlen
            units = tables.units() if self.share_scans else None
            bp_d = tables.bp_slice()
        else:
            lut = np.stack([tb['lut'] for tb in tables]).astype(np.float32)
            bias = np.array([tb['bias'] for tb in tables], dtype=np.float32)
            mlen = np.array([tb['m'] for tb in tables], dtype=np.int32)
            nfilt = np.array([tb['n_filters'] for tb in tables], dtype=np.int32)
            qtab = np.stack([tb['qtab'] for tb in tables]).view(np.int32)
            qthr = np.stack([tb['qthr'] for tb in tables]).view(np.int32)
            lut_d, bias_d, mlen_d, nfilt_d, qtab_d, qthr_d = (self._to_dev(lut), self._to_dev(bias), self._to_dev(mlen),
                                                              self._to_dev(nfilt), self._to_dev(qtab), self._to_dev(qthr))
            units = scan_units(tables) if self.share_scans else None
            bp_d = self._bp_tables(lut_d, bias_d, mlen_d, nfilt_d, T)
        if units is not None and len(units) > 2:
            # longest motifs first and similar lengths side by side: the two warps of a CTA then finish together and the
            # last wave of CTAs is the cheapest one
            units = units[np.argsort(-mlen[units[:, 0]], kind='stable')]
        units_d = self._to_dev(units) if units is not None else None
        n_units = len(units) if units is not None else 0
        self.last_units = n_units if units is not None else T
        self._last_units_arr = units
        self._units_run = getattr(self, '_units_run', 0) + self.last_units
        x_planes = entries = entry_count = row_nnz = None
        entry_cap = 0
        if path == 'tensor':
            x_bytes = int(_lib.lib.mepp_x_planes_bytes(m_rows, self.n_pad))
            x_planes = self._buf('x_planes', x_bytes, torch.uint8, zero=True)
        else:
            dens = max(SPARSE_ENTRY_DENSITY, 2.0 * getattr(self, '_density_est', 0.0))
            entry_cap = -(-max(MIN_ENTRY_CAP, int(T * N * L * dens)) // _lib.ENTRY_LISTS) * _lib.ENTRY_LISTS
            # (the match-driven profile behind the tensor-core scan never reads the entry list: not allocated then)
            tc_matches = (use_matches and use_lean and self.tc_scan and getattr(self, 'onehot', None) is not None
                          and not (self.bitpar and self.planes_T is not None))
            entries = None if tc_matches else self._buf(f'entries{slot}', entry_cap * 16, torch.uint8)
            entry_count = self._buf(f'entry_count{slot}', (16 * _lib.ENTRY_LISTS,), torch.int64)
            row_nnz = self._buf(f'row_nnz{slot}', (m_rows + 1,), torch.int32)
        rowstats = self._buf(f'rowstats{slot}', (m_rows, 3), torch.float64)
        count = self._buf(f'count{slot}', (T, L), torch.int32)
        # buffers that travel to the host are per staging slot: their copies run on a side stream while the next
        # group's kernels already write the other slot
        density = self._buf(f'density{slot}', (T, self.n_pad), torch.int32)
        hit_count = self._buf(f'hit_count{slot}', (1,), torch.int64)
        # the match list is always kept: it is also what restores the all-zero state of x_planes after the GEMM
        cap = int(hits_cap) if hits_cap else max(1 << 20, int(T * N * L * 2e-3))
        hits_d = self._buf(f'hits{slot}', cap * 16, torch.uint8)
        ev = self._mark(None, None)
        # tiled operand: per-task power-of-two scale of the fp16 hi + lo pieces (mepp_x_scales), undone by the GEMM epilogue
        xscale = None
        if path == 'tensor' and self.fuse_correlation:
            xscale = self._buf(f'xscale{slot}', (T,), torch.float32)
            _lib.check(_lib.lib.mepp_x_scales(_ptr(lut_d), _ptr(bias_d), _ptr(mlen_d), _ptr(nfilt_d), T, _ptr(xscale), st),
                       "x_scales")
            self.launches += 1
        scan_args = (_ptr(self.sym_T), _ptr(self.zhat), N, L, T, int(margin),
                     _ptr(lut_d), _ptr(bias_d), _ptr(mlen_d), _ptr(nfilt_d), _ptr(qtab_d), _ptr(qthr_d),
                     _ptr(units_d), n_units, int(mlen.max()), _ptr(x_planes), _ptr(rowstats), _ptr(count),
                     _ptr(density), _ptr(hits_d), cap, _ptr(hit_count), _ptr(entries), entry_cap,
                     _ptr(entry_count), _ptr(row_nnz))
        bitpar_on = bp_d is not None and self.planes_T is not None      # (the experimental first stage keeps its own path)
        if use_lean and self.tc_scan and path == 'sparse' and getattr(self, 'onehot', None) is not None and not bitpar_on:
            buckets = self._buf(f'buckets{slot}', int(_lib.lib.mepp_scan_lean_bucket_bytes(T, N)), torch.uint8)
            bucket_count = self._buf(f'bucket_count{slot}', (T * (self.n_pad // 32),), torch.int32)
            wbytes = int(_lib.lib.mepp_scan_tc_work_bytes(N, L, self.last_units))
            tc_work = self._buf('tc_work', wbytes, torch.uint8)
            units_h = np.ascontiguousarray(units, dtype=np.int32) if units is not None else None
            mlen_h = np.ascontiguousarray(mlen, dtype=np.int32)
            dbg = getattr(self, '_tc_dbg', None)
            _lib.check(_lib.lib.mepp_scan_tc(
                _ptr(self.sym_T), _ptr(self.onehot), _ptr(self.zhat), N, L, T, int(margin), _ptr(lut_d), _ptr(bias_d),
                _ptr(mlen_d), _ptr(nfilt_d), _ptr(units_d), units_h.ctypes.data if units_h is not None else None, n_units,
                mlen_h.ctypes.data, _ptr(rowstats), _ptr(count), _ptr(density), _ptr(hits_d), cap, _ptr(hit_count),
                _ptr(entries), entry_cap, _ptr(entry_count), _ptr(row_nnz), _ptr(buckets),
                _ptr(bucket_count), _ptr(tc_work), wbytes, _ptr(dbg), int(dbg.shape[1]) if dbg is not None else 0, st),
                "scan_tc")
            self.launches += 4
            self.last_scan = 'tc'
        elif use_lean and path == 'sparse' and not bitpar_on and L <= 65535:
            self.last_scan = 'lean'
            buckets = self._buf(f'buckets{slot}', int(_lib.lib.mepp_scan_lean_bucket_bytes(T, N)), torch.uint8)
            bucket_count = self._buf(f'bucket_count{slot}', (T * (self.n_pad // 32),), torch.int32)
            lean_args = scan_args[:15] + scan_args[16:] + (_ptr(buckets), _ptr(bucket_count))   # (no tiled operand)
            _lib.check(_lib.lib.mepp_scan_lean(*lean_args, st), "scan_lean")
            self.launches += 2
        elif bitpar_on:
            # units split by class on the device; the bit-parallel units and the pair-table units run as two launches
            bp_scratch = self._buf(f'bp_scratch{slot}', (4 * self.last_units + 2,), torch.int32)
            _lib.check(_lib.lib.mepp_scan_bitpar(*scan_args, _ptr(self.planes_T), _ptr(bp_d), _ptr(bp_scratch), _ptr(xscale),
                                                 st), "scan_bitpar")
            self.launches += 3
            self.last_bp = (bp_d, bp_scratch, self.last_units)
        else:
            _lib.check(_lib.lib.mepp_scan(*scan_args, _ptr(xscale), st), "scan")
            self.launches += 1
            self.last_scan = 'full'
        if path == 'tensor':
            # the scan stored only the non-zero pieces; the operand goes back to all-zero before its next use
            self._x_pending = dict(x=x_planes, hits=hits_d, cap=cap, hit_count=hit_count, N=N, L=L, T=T,
                                   margin=int(margin))
        ev = self._mark(ev, 'scan')
        # Everything after the scan runs as one chain.  With overlap it goes to a high-priority stream, so that the
        # next group's scan (enqueued on the compute stream as soon as this call returns) shares the SMs with it: the
        # scan is bound by integer issue and moves no memory, the chain is bound by L2 / HBM traffic.
        overlap = (self.overlap and path == 'sparse' and self.fuse_correlation and not return_hits and not return_r)
        compute_stream = torch.cuda.current_stream(self.device)
        if overlap:
            # (the score operand and its constants are built by the first chain, on the chain's stream: they wait for the
            # upload of the permutation indices, which must not hold up the scans of the later groups)
            if self._post_stream is None:
                self._post_stream = torch.cuda.Stream(device=self.device, priority=-1)
            scanned = torch.cuda.Event()
            scanned.record(compute_stream)
        elif self._post_done:
            torch.cuda.synchronize(self.device)   # a serial group after overlapped ones shares their scratch buffers
            self._post_done.clear()

        def chain():
            nonlocal ev
            st = self._stream()
            # tcgen05 GEMM with the correlation fused into its epilogue: r (m_rows, ldS) float32, S never reaches HBM
            r = self._buf('r', (m_rows, self.ldS), torch.float32)
            if path == 'sparse' and use_matches:
                # one gathered row of Y per MATCH (the blur commutes with the contraction), not per blurred non-zero
                work = self._buf('match_work', int(_lib.lib.mepp_profile_matches_work_bytes(m_rows, cap, C_)), torch.uint8)
                _lib.check(_lib.lib.mepp_profile_matches(_ptr(hits_d), _ptr(hit_count), cap, _ptr(count), T, L, int(margin),
                                                         _ptr(self._yrows()), self.n_pad, _ptr(rowstats),
                                                         _ptr(self._colconst(sparse=True)),
                                                         _ptr(r), self.ldS, C_, _ptr(work), st), "profile_matches")
                self.launches += 3
                nnz_d = work[m_rows * 4:m_rows * 4 + 4].view(torch.int32)     # offsets[m_rows] = matches of the group
                ev = self._mark(ev, 'gemm')
            elif path == 'sparse':
                work = self._buf('sparse_work', int(_lib.lib.mepp_profile_sparse_work_bytes(m_rows, entry_cap)), torch.uint8)
                _lib.check(_lib.lib.mepp_profile_sparse(_ptr(entries), _ptr(entry_count), entry_cap, _ptr(row_nnz),
                                                        _ptr(self._yrows()), _ptr(rowstats),
                                                        _ptr(self._colconst(sparse=True)), _ptr(r),
                                                        self.ldS, m_rows, C_, _ptr(work), st), "profile_sparse")
                self.launches += 3
                nnz_d = work[m_rows * 4:m_rows * 4 + 4].view(torch.int32)     # offsets[m_rows] = non-zeros of X
                ev = self._mark(ev, 'gemm')
            elif self.fuse_correlation:
                _lib.check(_lib.lib.mepp_profile_corr_gemm(_ptr(x_planes), _ptr(self.y_planes), _ptr(rowstats),
                                                           _ptr(self._colconst()), _ptr(r), m_rows, C_, self.n_pad,
                                                           _ptr(xscale), L, st), "profile_corr_gemm")
                self.launches += 2
                ev = self._mark(ev, 'gemm')
            else:
                S = self._buf('S', (m_rows, self.ldS), torch.float32)
                _lib.check(_lib.lib.mepp_profile_gemm(_ptr(x_planes), _ptr(self.y_planes), _ptr(S), m_rows, C_, self.n_pad,
                                                      st), "profile_gemm")
                ev = self._mark(ev, 'gemm')
                zsum = self._zsum()
                _lib.check(_lib.lib.mepp_correlation(_ptr(S), self.ldS, _ptr(rowstats), _ptr(self.col_syz), _ptr(self.ysum),
                                                     zsum.ctypes.data_as(C.POINTER(C.c_double)), N, m_rows, C_,
                                                     1 if self.use_gc else 0, _ptr(r), self.ldS, st), "correlation")
                self.launches += 3 + (m_rows + 65534) // 65535 - 1
                ev = self._mark(ev, 'correlation')
            dev = {}
            if P > 0:
                q = stats_host.percentile_q(ci)
                k_lo, _ = stats_host.virtual_index(P, q[0])
                k_hi, _ = stats_host.virtual_index(P, q[1])
                ks_lo, _ = stats_host.virtual_index(L * P, q[0])
                ks_hi, _ = stats_host.virtual_index(L * P, q[1])
                dev['full_os'] = self._buf('full_os', (T, L, 4), torch.float32)
                dev['full_cnt'] = self._buf('full_cnt', (T, L), torch.int32)
                dev['semi_os'] = self._buf('semi_os', (T, 4), torch.float32)
                dev['semi_cnt'] = self._buf('semi_cnt', (T, L), torch.int64)
                dev['integral'] = self._buf('integral', (T, C_), torch.float64)
                work = self._buf('stats_work', int(_lib.lib.mepp_perm_stats_work_bytes(T, L, C_)), torch.uint8)
                _lib.check(_lib.lib.mepp_perm_stats(_ptr(r), self.ldS, T, L, C_, k_lo, k_hi, ks_lo, ks_hi, _ptr(dev['full_os']),
                                                    _ptr(dev['full_cnt']), _ptr(dev['semi_os']), _ptr(dev['semi_cnt']),
                                                    _ptr(dev['integral']), _ptr(work), st), "perm_stats")
                self.launches += 9
                ev = self._mark(ev, 'perm_stats')
            # K4: every output column of the group on the device; the host only copies the block into the result table
            out_d = self._buf(f'finalized{slot}', int(_lib.lib.mepp_finalize_out_bytes(T, L)), torch.uint8)
            if P > 0:
                _, g_lo = stats_host.virtual_index(P, q[0])
                _, g_hi = stats_host.virtual_index(P, q[1])
                _, gs_lo = stats_host.virtual_index(L * P, q[0])
                _, gs_hi = stats_host.virtual_index(L * P, q[1])
                fin = (dev['full_os'], dev['full_cnt'], dev['semi_os'], dev['semi_cnt'], dev['integral'])
            else:
                k_lo = k_hi = 0
                g_lo = g_hi = gs_lo = gs_hi = 0.0
                fin = (None,) * 5
            _lib.check(_lib.lib.mepp_finalize_positions(_ptr(r), self.ldS, T, L, C_, *[_ptr(x) for x in fin], _ptr(count), N,
                                                        int(margin), k_lo, k_hi, g_lo, g_hi, gs_lo, gs_hi,
                                                        float(stats_host.z_margin_f32(N, ci)), _ptr(out_d), st),
                       "finalize_positions")
            self.launches += 2 if P > 0 else 1
            ev = self._mark(ev, 'finalize')
            dev = {'finalized': out_d}
            if L <= 65535:
                d16 = self._buf(f'density16_{slot}', (T, self.n_pad), torch.int16)
                _lib.check(_lib.lib.mepp_density_u16(_ptr(density), T * self.n_pad, _ptr(d16), st), "density_u16")
                self.launches += 1
                dev['density'] = d16
            else:
                dev['density'] = density
            dev['hit_count'] = hit_count
            if heatmap is not None:
                # block maxima of the clipped score matrix on the figure's pixel grid, from the match list of the group
                from . import heatmap as _hm
                bh, bw = _hm.get_aspect_tensor((N, L), heatmap[0], heatmap[1])
                H, W = -(-N // bh), -(-L // bw)
                heat = self._buf(f'heat{slot}', (T, H, W), torch.float32)
                hwork = self._buf('heat_work', int(_lib.lib.mepp_heatmap_work_bytes(T)), torch.uint8)
                sig = -1.0 if heatmap[2] is None else float(heatmap[2])
                _lib.check(_lib.lib.mepp_heatmap(_ptr(hits_d), _ptr(hit_count), cap, T, N, L, bh, bw, sig, _ptr(heat),
                                                 _ptr(hwork), st), "heatmap")
                self.launches += 3
                dev['heatmap'] = heat
            if path == 'sparse':
                dev['entry_count'] = entry_count
                dev['nnz'] = nnz_d
            if return_r:
                dev['r'] = r.view(T, L, self.ldS)[:, :, :C_].contiguous()
            host = {}
            # device -> host on a side stream: the next group's kernels do not wait for the copies
            compute = torch.cuda.current_stream(self.device)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            ready = torch.cuda.Event()
            ready.record(compute)
            self._copy_stream.wait_event(ready)
            with torch.cuda.stream(self._copy_stream):
                for k, t in dev.items():
                    h = self._pinned(k, slot, t.shape, t.dtype)
                    h.copy_(t, non_blocking=True)
                    host[k] = h
                done = torch.cuda.Event()
                done.record(self._copy_stream)
            return dict(host=host, done=done, T=T, hits_d=hits_d, cap=cap, return_hits=return_hits, path=path,
                        use_matches=use_matches,
                        entry_cap=entry_cap, slot=slot, tables=tables, return_r=return_r, hits_cap=hits_cap,
                        heatmap=heatmap,
                        dev=dev, chain_done=ready)    # (keeps the copied tensors alive until the side stream is done with them)

        if not overlap:
            return chain()
        with torch.cuda.stream(self._post_stream):
            self._post_stream.wait_event(scanned)
            out = chain()
            self._post_done[slot] = out['chain_done']
        return out

    def _collect_group(self, pend, margin, ci, store, density_all):
        """Wait for one group's D2H copies and finalise its statistics on the host, straight into the
        preallocated output store (rows t0 .. t0+T)."""
        pend['done'].synchronize()
        import time as _time
        self._last_wait_end = _time.perf_counter()
        h = {k: v.numpy() for k, v in pend['host'].items()}
        t0, T = pend['t0'], pend['T']
        if pend['path'] == 'sparse':
            list_full = int(h['entry_count'][::16].max()) > pend['entry_cap'] // _lib.ENTRY_LISTS
            lean_full = int(h['entry_count'][1]) != 0      # a bucket (or a candidate list) of the lean scan overflowed
            hits_full = pend.get('use_matches') and int(h['hit_count'][0]) > pend['cap']
            if list_full or lean_full or hits_full:
                # a lean overflow (a run of matches inside 32 sequences) is redone by the full-form scan, still on the
                # sparse path (from the blurred entries when the match list itself was too short); only an X too dense
                # for the entry list goes to the tensor-core path
                self.fallbacks.append('tensor' if list_full else 'full-scan' if lean_full else 'entries')
                redo = self._launch_group(pend['tables'], margin, pend['return_hits'], pend['return_r'], pend['hits_cap'],
                                          ci, pend['slot'], path='tensor' if list_full else 'sparse',
                                          lean=False if lean_full or list_full else None,
                                          matches=False if hits_full else None, heatmap=pend.get('heatmap'))
                redo['t0'] = t0
                redo['call'] = pend.get('call')
                if redo['call'] is not None:
                    redo['call'].paths.append(redo['path'])
                return self._collect_group(redo, margin, ci, store, density_all)
        # host side of a group = copies of the finalised block and of the densities into the result table;
        # split over a few threads (memcpy drops the GIL) so that the exposed tail of the last group stays short
        N = self.N

        def job(a, b):
            stats_host.store_finalized(store, t0, h['finalized'], T, self.L, a, b)
            density_all[t0 + a:t0 + b] = h['density'].view(density_all.dtype)[a:b, :N]

        n_chunks = max(1, min(self.copy_threads, T // 8))
        bounds = [T * i // n_chunks for i in range(n_chunks + 1)]
        if n_chunks == 1:
            job(0, T)
        else:
            for f in [_copy_pool(self.copy_threads).submit(job, bounds[i], bounds[i + 1]) for i in range(n_chunks)]:
                f.result()
        out = dict(T=T, t0=t0)
        out['entries'] = int(h['nnz'][0]) if 'nnz' in h else 0
        nh = int(h['hit_count'][0])
        out['hit_count'] = np.array([nh], dtype=np.int64)
        if pend['return_hits']:
            if nh > pend['cap']:
                raise RuntimeError(f"match list overflow: {nh} matches > capacity {pend['cap']}; pass hits_cap")
            out['hits'] = pend['hits_d'][:nh * 16].cpu().numpy().view(HIT_DTYPE).copy()
        if 'r' in h:
            out['r'] = h['r'].copy()
        if 'heatmap' in h:
            if nh > pend['cap']:
                raise RuntimeError(f"match list overflow: {nh} matches > capacity {pend['cap']}; pass hits_cap")
            out['heatmap'] = h['heatmap'].copy()
        return out
