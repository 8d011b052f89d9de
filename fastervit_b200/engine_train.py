"""Training plan: train-mode forward (batch-statistics BatchNorm, saved activations) and the backward pass
of FasterViT as static launch lists over libfvit_sm100.so.

The reference has no backward code — fv.py relies on autograd (train.py:879-896) — so this module *is*
the backward of SURVEY §8 row a16: per layer it emits the weight-gradient GEMM (both operands MN-major:
dW = dZ^T X, split-K with fp32 atomics into one flat gradient buffer), the data-gradient GEMM (B = the
forward's packed weight read as an MN-major operand, so no transposed copies exist), and the HBM-bound
reductions (bias / LayerNorm / BatchNorm / layer-scale / positional-MLP gradients).

Numerics: activation gradients are fp16 operands (fp32 along the residual stream) multiplied by a
power-of-two scale S chosen on the device from max|dlogits| (scal[0] = S, scal[1] = 1/S); a branch whose
layer scale gamma is tiny (1e-5 at init for fv3+) is additionally normalised by s_b = 2^-floor(log2 max|gamma|)
so no fp16 product underflows. Parameter gradients are fp32 and exact in scale.
"""
from __future__ import annotations

import ctypes as C

import os

import torch

from . import lib as L
from .engine import Plan, _ru
from .engine_train_conv import ConvPartEmitters
from .engine_train_levels import TokenLevelEmitters


class TrainPlan(Plan, ConvPartEmitters, TokenLevelEmitters):
    def __init__(self, model, B: int, H: int, W: int, device):
        self.bwd_ops: list[tuple] = []
        self.fwd_zero: list[torch.Tensor] = []   # accumulators zeroed before every forward (BN statistics)
        self.saved_levels: list[dict] = []
        super().__init__(model, B, H, W, True, device)

    # ------------------------------------------------------------------------------ infrastructure
    def _setup_train(self) -> None:
        m, nb = self.model, self.bufs
        self.params = [p for p in m.parameters()]
        offs, n = {}, 0
        for p in self.params:
            offs[id(p)] = n
            n += _ru(p.numel(), 64)            # 256-byte aligned slices
        self.gflat = nb.new("grad.flat", (n,), torch.float32)
        self._goff = offs
        self.scratch_n = 0
        self._scratch_req: list[tuple[str, int]] = []
        self._bwd_side: set[int] = set()     # backward launches that run as a side branch (engine_train_levels.py)
        self._bwd_side_wgrad: set[int] = set()    # weight-gradient launches: side branch on a single GPU only (below)
        self._side_reads: dict[int, tuple] = {}   # side launch -> shared gradient buffers it reads (WAR hazards)
        # device scalars: [0] S, [1] 1/S, [2] 1.0, then per-branch (s_b, 1/s_b, 1/(S s_b))
        self._nscal = 3
        self._branch_slots: list[tuple[int, object]] = []

    def G(self, p: torch.Tensor) -> int:
        """device pointer of the flat-buffer gradient slice of parameter p"""
        return self.gflat.data_ptr() + 4 * self._goff[id(p)]

    def grad_views(self) -> list[torch.Tensor]:
        """Per-parameter views of a *copy* of the flat gradient buffer: autograd's AccumulateGrad may keep
        the tensors it is handed, and the flat buffer itself is reused (zeroed) by the next backward."""
        flat = self.gflat.clone()
        return [flat[self._goff[id(p)]: self._goff[id(p)] + p.numel()].view_as(p) for p in self.params]

    def _scratch(self, name: str, n: int) -> int:
        """reserve n zero-initialised fp32 (re-zeroed at the start of every backward); returns offset"""
        off = self.scratch_n
        self.scratch_n += _ru(n, 64)
        self._scratch_req.append((name, off))
        return off

    def _branch(self, gamma) -> dict:
        """device-scalar pointers of a residual branch: dgrad alpha (1/s_b) and wgrad alpha (1/(S s_b))"""
        if not isinstance(gamma, torch.Tensor):
            return dict(gamma=None, s=None, inv_s=("scal", 2), w_alpha=("scal", 1))
        slot = self._nscal
        self._nscal += 3
        self._branch_slots.append((slot, gamma))
        return dict(gamma=gamma, s=("scal", slot), inv_s=("scal", slot + 1), w_alpha=("scal", slot + 2))

    def _finish_train(self) -> None:
        nb = self.bufs
        self.scal = nb.new("grad.scalars", (_ru(self._nscal, 16),), torch.float32)
        self.scal[2] = 1.0
        self.scr = nb.new("grad.scratch", (max(self.scratch_n, 64),), torch.float32)
        sp = self.scal.data_ptr()
        for slot, gamma in self._branch_slots:
            self._op(self.prep_ops, "fvit_pow2_norm", gamma.data_ptr(), gamma.numel(), sp + 4 * slot)
        # resolve symbolic pointers ("scal", i) / ("scr", off) in the op lists
        def res(v):
            if isinstance(v, tuple) and len(v) == 2 and v[0] == "scal":
                return sp + 4 * v[1]
            if isinstance(v, tuple) and len(v) == 2 and v[0] == "scr":
                return self.scr.data_ptr() + 4 * v[1]
            return v
        for lst in (self.ops, self.bwd_ops, self.prep_ops):
            for i, (fn, args, name) in enumerate(lst):
                if isinstance(args, tuple):
                    lst[i] = (fn, tuple(res(a) for a in args), name)
        for g in self._gemm_keep:
            for f in ("alpha_ptr", "out_f32", "col_sum", "col_sumsq", "out_colsum", "out_colsum_alpha"):
                v = getattr(g, "_sym_" + f, None)
                if v is not None:
                    setattr(g, f, res(v))
        # per-backward prologue: zero accumulators, pick the gradient scale, per-branch wgrad alphas
        pro: list[tuple] = [("zero", self.gflat, "memset"), ("zero", self.scr, "memset")]
        for lv in self.lv:
            pro.append(("zero", lv["g"], "memset"))
        self._bwd_prologue = pro
        self._branch_alpha_ops = []
        for slot, _ in self._branch_slots:
            self._branch_alpha_ops.append((self.lib.fvit_vec_mul, (sp + 4, 0, sp + 4 * (slot + 1), 0, sp + 4 * (slot + 2), 1),
                                           "fvit_vec_mul"))

    # symbolic-pointer aware GEMM emitter for the backward list
    def _bgemm(self, *, a, a_rows, lda, b, b_rows, ldb, m, n, kc, a_mn=False, b_mn=False, taps=None, a_planes=1,
               a_plane_stride=0, b_row_off=0, b_taps=None, split_k=1, alpha_ptr=None, act=L.ACT_NONE, aux=None, ld_aux=0,
               out_f32=None, ld_o32=0, out_f16=None, ld_o16=0, row_map=None, resid=None, ld_resid=0, target=None,
               flops=None, out_colsum=None, out_colsum_alpha=None, aux_scale=None, aux_shift=None) -> None:
        g = L.GemmArgs()
        g.a, g.a_rows, g.lda, g.a_plane_stride, g.a_planes, g.a_mn_major = a, a_rows, lda, a_plane_stride, a_planes, int(a_mn)
        g.b, g.b_rows, g.ldb, g.b_mn_major = b, b_rows, ldb, int(b_mn)
        g.m, g.n, g.kc = m, n, kc
        taps = taps or [(0, 0)]
        g.ntaps = len(taps)
        for i, (s, p) in enumerate(taps):
            g.tap_shift[i], g.tap_plane[i] = s, p
        g.b_row_off = b_row_off
        if b_taps is not None:   # conv weight gradient: taps folded into N (include/fvit.h: b_ntaps)
            assert len(taps) == 1 and a_mn and b_mn
            g.b_ntaps = len(b_taps)
            for i, sft in enumerate(b_taps):
                g.tap_shift[i] = sft
        g.split_k, g.alpha, g.act = split_k, 1.0, act
        g.aux, g.ld_aux, g.row_map = aux, ld_aux, row_map
        g.aux_scale, g.aux_shift = aux_scale, aux_shift
        g.resid, g.ld_resid = resid, ld_resid
        g.out_f16, g.ld_out_f16, g.ld_out_f32 = out_f16, ld_o16, ld_o32
        for f, v in (("alpha_ptr", alpha_ptr), ("out_f32", out_f32), ("out_colsum", out_colsum),
                     ("out_colsum_alpha", out_colsum_alpha)):
            if isinstance(v, tuple):
                setattr(g, "_sym_" + f, v)
            elif v is not None:
                setattr(g, f, v)
        self._gemm_keep.append(g)
        lst = self.bwd_ops if target is None else target
        lst.append((self.lib.fvit_gemm, (C.byref(g),), "fvit_gemm"))
        if lst is self.bwd_ops:
            self.bwd_flops[len(lst) - 1] = flops if flops is not None else 2.0 * m * n * kc * len(taps)

    # ------------------------------------------------------------------------------ weight gradients as a side branch
    # A weight-gradient GEMM reads an activation gradient (a per-level buffer every block reuses) and a saved
    # activation, and adds into parameter-gradient memory nobody reads before the next bucket point: a leaf. As a
    # side-branch launch it runs beside the data-gradient chain -- two persistent grids share the SMs, each filling
    # the other's tail waves. The one hazard is write-after-read on the reused gradient buffers: the launch that next
    # overwrites such a buffer first waits for the event behind its last side-branch reader (`wait_side`).
    def _side_from(self, first: int, *reads: int) -> None:
        """make backward launches [first, now) side-branch launches that read the reusable buffers `reads`"""
        if not self.wgrad_side:
            return
        for i in range(first, len(self.bwd_ops)):
            self._bwd_side_wgrad.add(i)
            self._side_reads[i] = tuple(reads)

    def _side_set(self, data_parallel: bool):
        """Launches that go to the side branch. Under a gradient all-reduce the weight-gradient GEMMs stay there too: a
        second persistent grid does compete with NCCL's CTAs for the SMs the capped launches leave free, but measured on
        2 x B200 (profiles/r02v_ddp_n2.txt) the step is 75.1 ms with them against 76.8 ms without (73.0 ms on one GPU of
        the same box). FVIT_WGRAD_SIDE_DDP=0 keeps them on the main branch (A/B switch)."""
        if not self.side_branches:
            return None
        if data_parallel and os.environ.get("FVIT_WGRAD_SIDE_DDP", "1") != "1":
            return self._bwd_side
        return self._bwd_side | self._bwd_side_wgrad

    def _before_write(self, *ptrs: int) -> None:
        """the next backward launch overwrites `ptrs`: order it behind their side-branch readers"""
        if self.wgrad_side:
            self.bwd_ops.append(("wait_side", tuple(ptrs), "wait_side"))

    def _split_k(self, m: int, n: int, k_rows: int, ntaps: int = 1) -> int:
        tiles = ((m + 127) // 128) * ((n + 255) // 256) * ntaps
        kb = (k_rows + 63) // 64
        best, best_cost = 1, float("inf")
        for sk in range(1, max(1, min(64, kb // 8)) + 1):
            # time model (us): waves of CTAs x (K blocks of one unit x 0.26 us per 128x256x64 block + epilogue);
            # the split-K epilogue is a pass of fp32 vector reductions into L2 (~10 us per tile, measured)
            cost = -(-tiles * sk // 148) * (kb / sk * 0.26 + (10.0 if sk > 1 else 3.0))
            if cost < best_cost - 1e-9:
                best, best_cost = sk, cost
        return best

    # ------------------------------------------------------------------------------ linear layer backward
    def _linear_bwd(self, *, lin, w16, ldw, x16, ldx, dz16, lddz, rows, n_out, k_in, br, gW=None, gW_ld=None,
                    dx16=None, lddx=0, dx_act=L.ACT_NONE, dx_aux=None, ld_aux=0, dx_alpha=None, bias_to=None,
                    flops_k=None, want_dgrad=True, bias_done=False, dx_colsum=None, gW_row_map=None) -> None:
        """dW[n_out, k_in] += alpha_w * dz16^T @ x16 ; db += alpha_w * colsum(dz16) ; dx16 = alpha_d * dz16 @ W."""
        gW = self.G(lin.weight) if gW is None else gW
        gW_ld = k_in if gW_ld is None else gW_ld
        fk = k_in if flops_k is None else flops_k
        n0 = len(self.bwd_ops)
        self._bgemm(a=dz16, a_rows=rows, lda=lddz, a_mn=True, b=x16, b_rows=rows, ldb=ldx, b_mn=True, m=n_out, n=k_in,
                    kc=rows, split_k=self._split_k(n_out, k_in, rows), alpha_ptr=br["w_alpha"], out_f32=gW,
                    ld_o32=gW_ld, row_map=gW_row_map, flops=2.0 * rows * n_out * fk)
        if (lin.bias is not None or bias_to is not None) and not bias_done:
            dst = bias_to if bias_to is not None else self.G(lin.bias)
            self._op(self.bwd_ops, "fvit_colsum", dz16, 1, lddz, None, None, 0, rows, n_out, None, br["w_alpha"], dst, None)
        self._side_from(n0, dz16)   # (weight gradient + bias column sums: both only read dz16)
        if want_dgrad:
            self._before_write(dx16)
            self._bgemm(a=dz16, a_rows=rows, lda=lddz, b=w16, b_rows=n_out, ldb=ldw, b_mn=True, m=rows, n=k_in, kc=n_out,
                        alpha_ptr=dx_alpha if dx_alpha is not None else br["inv_s"], act=dx_act, aux=dx_aux,
                        ld_aux=ld_aux, out_f16=dx16, ld_o16=lddx, flops=2.0 * rows * n_out * fk,
                        out_colsum=dx_colsum[0] if dx_colsum else None,
                        out_colsum_alpha=dx_colsum[1] if dx_colsum else None)

    # ------------------------------------------------------------------------------ plan construction
    def _build(self) -> None:
        m, B = self.model, self.B
        self.bwd_flops: dict[int, float] = {}
        self._setup_train()
        self._build_forward()
        self._build_backward()
        self._finish_train()

    # forward / backward emitters: the ConvPartEmitters / TokenLevelEmitters mixins (engine_train_conv.py, engine_train_levels.py)

    # ------------------------------------------------------------------------------ execution
    def run_forward(self, x: torch.Tensor) -> torch.Tensor:
        # saved activations / batch statistics / drop masks live in this plan's persistent buffers: a new forward
        # invalidates the state an earlier, not yet differentiated forward left behind (checked in backward)
        self.fwd_generation = getattr(self, "fwd_generation", 0) + 1
        L.bump_weights_epoch()   # BatchNorm running statistics are updated through raw pointers
        def body(xin):
            for t in self.fwd_zero:
                t.zero_()
            if self.drop_specs:
                self._gen_drop_masks()   # stochastic-depth masks of this step (torch RNG, like timm's DropPath)
            if self.prep_branches:
                self.run_ops_branches(self.prep_ops)
            else:
                self.run_ops(self.prep_ops, None)
            self.run_ops(self.ops, xin)
            for bn in self._bn_modules:
                bn.num_batches_tracked += 1
        if self.use_graphs and not getattr(self, "forced_drop_masks", None):
            xs_ = self.static_input(x)
            self.run_captured("train forward", lambda: body(xs_))
        else:
            body(x)
        return self.logits

    def _bwd_start(self, dlogits: torch.Tensor | None) -> None:
        if dlogits is not None:
            self._dlogits[:, :self._ncls].copy_(dlogits)
        self.run_ops(self._bwd_prologue, None)
        st = L.stream_ptr()
        rc = self.lib.fvit_grad_scale_init(self._dlogits.data_ptr(), self._dlogits.numel(), 64.0, self.scal.data_ptr(), st)
        if rc:
            raise L.FvitError(self.lib.fvit_last_error().decode())
        self.run_ops(self._branch_alpha_ops, None)

    def run_backward(self, dlogits: torch.Tensor) -> None:
        grp = getattr(self.model, "_grad_allreduce", None)
        if grp is None:
            self._dlogits[:, :self._ncls].copy_(dlogits)

            def body():
                self._bwd_start(None)
                self.run_ops(self.bwd_ops, None, side=self._side_set(False))
            self.run_captured("backward", body)
            return
        # data parallel: each bucket's all-reduce is issued the moment the launches that finish its gradients are
        # enqueued, so NCCL (its own stream) overlaps the rest of the backward pass. The launch list is cut at the
        # bucket points: every segment is one CUDA graph (like the single-GPU backward), the all-reduces in between are
        # ordinary NCCL calls -- collectives captured INTO a graph work too, but live graphs holding NCCL work make
        # destroy_process_group() hang at exit (seen on 2 x B200, r02j), so they stay outside.
        self._dlogits[:, :self._ncls].copy_(dlogits)
        red = GradBucketReducer(self.gflat, None if grp is True else grp)
        if getattr(self, "_ar_shadow", None) is None:
            self._ar_shadow = self._allreduce_shadow(red.world, red.measure_busbw())
        if getattr(self, "_bwd_segments", None) is None:
            segs, start = [], 0
            for i, op in enumerate(self.bwd_ops):
                if op[0] == "bucket":
                    segs.append((start, i, op[1]))
                    start = i + 1
            segs.append((start, len(self.bwd_ops), None))
            self._bwd_segments = segs
        self._ar_active = red
        try:
            for n, (lo, hi, bucket) in enumerate(self._bwd_segments):
                def body(lo=lo, hi=hi, first=(n == 0)):
                    if first:
                        self._bwd_start(None)
                    self._op_base = lo
                    try:
                        self.run_ops(self.bwd_ops[lo:hi], None, side=self._side_set(True))
                    finally:
                        self._op_base = 0
                if hi > lo or n == 0:   # (the list ends with a bucket point: nothing to capture after it)
                    self.run_captured(f"backward segment {n}", body)
                if bucket is not None:
                    red.reduce(*bucket)
        finally:
            self._ar_active = None
        red.finish(self.grad_buckets)

    def _allreduce_shadow(self, world: int, busbw: float | None = None) -> dict:
        """{backward op index: SM cap} for the launches that run while a bucket's all-reduce is in flight. NCCL's CTAs
        cannot co-reside with a GEMM CTA (227 KB of shared memory, one per SM): a full-width persistent grid launched
        into the reduction would run its last CTAs as a second wave, doubling that launch. The window is estimated at
        plan time: bucket bytes over the measured all-reduce bus bandwidth (B200_PROFILING.md: 725 GB/s at 8 ranks)
        against the launches' algorithmic FLOPs at the model's typical 700 TFLOP/s; those launches are capped to
        148 - FVIT_NCCL_CTAS SMs (default 16 = NCCL_MAX_CTAS set by enable_grad_allreduce callers such as bench.py)."""
        import os
        reserve = int(os.environ.get("FVIT_NCCL_CTAS", "16"))
        if world <= 1 or reserve <= 0:
            return {}
        cap = max(8, 148 - reserve)
        shadow: dict[int, int] = {}
        # all-reduce bus bandwidth with NCCL_MAX_CTAS = 16: 8 ranks ~600 GB/s (NVLS), fewer ranks ring over fewer links
        # (timeline at 2 ranks, profiles/r02h_ddp_timeline.json: the 790 MB level-2 bucket shadows ~6 ms of launches)
        # -- measured once per plan on the idle fabric (GradBucketReducer.measure_busbw); the table is the fallback
        if not busbw or busbw <= 0:
            busbw = {2: 200e9, 3: 300e9, 4: 400e9}.get(world, 600e9)
        self.ar_busbw = busbw
        for i, op in enumerate(self.bwd_ops):
            if op[0] != "bucket":
                continue
            lo, hi = op[1]
            t_need = 1.5 * 2.0 * (world - 1) / world * (hi - lo) * 4 / busbw   # seconds, with margin
            t = 0.0
            for j in range(i + 1, len(self.bwd_ops)):
                if t >= t_need:
                    break
                shadow[j] = cap
                t += max(self.bwd_flops.get(j, 0.0) / 700e12, 8e-6)
        return shadow

    def profile(self, x: torch.Tensor) -> list[dict]:
        """Per-launch CUDA-event timing of one training step (forward list, then backward list with the
        gradient of sum(logits)/B as a stand-in loss gradient)."""
        def timed(name, fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            torch.cuda._sleep(int(2e8))
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            return dict(name=name, ms=e0.elapsed_time(e1), flops=0.0, phase="step")

        for t in self.fwd_zero:
            t.zero_()
        # per-step work around the two launch lists: weight re-packing (parameters change every optimizer step),
        # accumulator clears, the copy that hands the flat gradient buffer to autograd
        out = [timed("prep: repack weights to fp16 operands", lambda: self.run_ops(self.prep_ops, None)),
               timed("backward prologue: clear gradient buffers", lambda: self.run_ops(self._bwd_prologue, None)),
               timed("gradient hand-off copy", lambda: self.grad_views())]
        for ops, flops, tag in ((self.ops, self.op_flops, ""), (self.bwd_ops, self.bwd_flops, "")):
            if ops is self.bwd_ops:
                self._bwd_start(torch.full((self.B, self._ncls), 1.0 / self.B, device=self._dlogits.device))
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)]
            torch.cuda.synchronize()
            torch.cuda._sleep(int(8e8))
            evs[0].record()
            for i, op in enumerate(ops):
                self.run_ops([op], x)
                evs[i + 1].record()
            torch.cuda.synchronize()
            out += [dict(name=op[2], ms=evs[i].elapsed_time(evs[i + 1]), flops=flops.get(i, 0.0),
                         phase="bwd" if ops is self.bwd_ops else "fwd", **self._op_desc(op))
                    for i, op in enumerate(ops) if op[0] not in ("wait_side", "bucket")]   # (markers launch nothing)
        return out


class GradBucketReducer:
    """Asynchronous mean all-reduce of slices of the flat gradient buffer (train.py:542-551's DDP, restated for
    a single autograd node): `reduce(lo, hi)` is called from the backward launch list at the points where the
    slice [lo, hi) is final; `finish` reduces whatever was not covered and waits for all of them."""

    def __init__(self, flat: torch.Tensor, group=None):
        import torch.distributed as dist
        if not dist.is_available() or not dist.is_initialized():
            raise L.FvitError("enable_grad_allreduce() needs an initialised torch.distributed process group")
        self.dist, self.flat, self.group = dist, flat, group
        self.world = dist.get_world_size(group)
        self.avg = dist.get_backend(group) == "nccl"   # gloo has no AVG
        self.pending: list = []
        self.done: list[tuple[int, int]] = []

    def reduce(self, lo: int, hi: int) -> None:
        if self.world == 1 or hi <= lo:
            return
        op = self.dist.ReduceOp.AVG if self.avg else self.dist.ReduceOp.SUM
        self.pending.append(self.dist.all_reduce(self.flat[lo:hi], op=op, group=self.group, async_op=True))
        self.done.append((lo, hi))

    def measure_busbw(self, nbytes: int = 64 << 20) -> float | None:
        """All-reduce bus bandwidth (bytes/s) of this group as configured (NCCL_MAX_CTAS, topology), timed once on a
        slice of the flat buffer before the backward pass zeroes it: sizes the SM cap windows of the launch list."""
        if self.world == 1 or not self.avg or not self.flat.is_cuda:
            return None
        n = min(self.flat.numel(), nbytes // 4)
        buf = self.flat[:n]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(3):   # two warm-ups (channel setup), one timed
            if i == 2:
                ev[0].record()
            self.dist.all_reduce(buf, op=self.dist.ReduceOp.AVG, group=self.group)
        ev[1].record()
        ev[1].synchronize()
        ms = ev[0].elapsed_time(ev[1])
        bw = torch.tensor([2.0 * (self.world - 1) / self.world * n * 4 / (ms * 1e-3)], device=self.flat.device)
        self.dist.all_reduce(bw, op=self.dist.ReduceOp.MIN, group=self.group)   # every rank plans the same windows
        return float(bw.item())

    def finish(self, buckets: list[tuple[int, int]]) -> None:
        for lo, hi in buckets:
            if (lo, hi) not in self.done:
                self.reduce(lo, hi)
        for w in self.pending:
            w.wait()
        if not self.avg and self.world > 1:
            self.flat.div_(self.world)


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Data parallel (train.py:542-551): the ONE exchange of the path — the flat gradient buffer is averaged
    over the ranks with a single all-reduce (NCCL over NVLink/NVSwitch on the B200 box; gloo in the CPU
    tests). BatchNorm statistics stay per-GPU like the reference (sync_bn: false)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise L.FvitError("enable_grad_allreduce() needs an initialised torch.distributed process group")
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:  # gloo has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat


class _FasterViTFunction(torch.autograd.Function):
    """One autograd node for the whole network: forward runs the train plan, backward runs the backward
    plan and hands out views of the flat gradient buffer (one per nn.Parameter, in .parameters() order)."""

    @staticmethod
    def forward(ctx, plan: TrainPlan, x: torch.Tensor, *params):
        if x.requires_grad:
            raise L.FvitError("the input requires grad, but the backward plan stops at the first convolution (no "
                              "gradient w.r.t. the image is computed); detach the input")
        ctx.plan = plan
        logits = plan.run_forward(x)
        ctx.generation = plan.fwd_generation
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        plan: TrainPlan = ctx.plan
        if ctx.generation != plan.fwd_generation:
            raise L.FvitError(
                "backward of a stale forward: another train-mode forward of the same input shape ran in between and "
                "overwrote the saved activations of this one (one autograd node per step: call backward before the "
                "next forward, or use different plans / model copies for multi-view losses)")
        plan.run_backward(dlogits.contiguous().float())
        return (None, None, *plan.grad_views())
