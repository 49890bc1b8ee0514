"""The train step of train.py:345-354 as one replayable unit (our public API for training / the benchmark):

    H2D(tokens, prefix) -> step clock (RNG + LR schedule) -> noise_injection -> forward -> masked CE -> backward
    -> [NCCL all-reduce of the flat gradient buffer] -> fused HF-AdamW (+ zero_grad)

Everything between the H2D copies and the optimizer is captured in CUDA graphs (shapes are static per config).
Data parallel: one process per GPU, captions sharded across ranks, ONE all-reduce per step over the
flat buffer [n_valid, loss_sum, 0, 0 | grads...]: CE gradients are sum-reduced per rank and divided by the GLOBAL
count of non-ignored targets inside the AdamW kernel, so N ranks reproduce the single-GPU global-batch mean exactly.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import ops
from ._lib import CapdecError


def push_plan(span, world: int, rank: int, shard: int):
    """Where rank `rank` sends the gradients [span[0], span[1]) (trainable-parameter coordinates) of a finished bucket:
    one entry per OTHER rank r whose slice [r * shard, (r + 1) * shard) the span touches -
    (r, first element, number of elements, element offset inside rank r's staging area).  A rank's staging area holds
    world - 1 slots of `shard` floats, one per source rank in rank order with the owner itself left out."""
    out = []
    lo, hi = span
    for r in range(world):
        if r == rank:
            continue
        a, b = max(lo, r * shard), min(hi, (r + 1) * shard)
        if b > a:
            slot = rank if rank < r else rank - 1
            out.append((r, a, b - a, slot * shard + (a - r * shard)))
    return out


class Trainer:
    def __init__(self, model, batch_size: int, seq_len: int = 40, lr: float = 2e-5, warmup_steps: int = 5000,
                 total_steps: int = 100000, noise_variance: float = 0.0, uniform_noise: bool = False,
                 dont_norm: bool = False, modality_offset: Optional[torch.Tensor] = None, betas=(0.9, 0.999),
                 eps: float = 1e-6, weight_decay: float = 0.0, use_cuda_graph: bool = True, process_group=None):
        self.model = model
        self.eng = model.engine()
        self.dev = self.eng.dev
        self.B, self.L = batch_size, seq_len
        self.lr, self.warmup, self.total = lr, warmup_steps, total_steps
        self.noise_variance, self.uniform_noise, self.dont_norm = noise_variance, uniform_noise, dont_norm
        self.offset = None
        if modality_offset is not None:
            self.offset = modality_offset.to(device=self.dev, dtype=torch.float32).reshape(-1).contiguous()
        self.betas, self.eps, self.wd = betas, eps, weight_decay
        self.train_gpt = model.gpt_trainable()
        fl = self.eng.flat
        self.n_train = fl.n_mapper + (fl.n_gpt if self.train_gpt else 0)
        lo, hi = fl.tail, fl.tail + self.n_train
        self.p_flat, self.g_flat = fl.params[lo:hi], fl.grads[lo:hi]
        self.m_flat, self.v_flat = torch.zeros_like(self.p_flat), torch.zeros_like(self.p_flat)
        self.reduce_buf = fl.grads[:hi]                      # [tail | trainable grads]: the single all-reduce payload
        self.tail = fl.grads[:4]
        self.step_dev = torch.zeros(1, device=self.dev)
        self.lr_dev = torch.zeros(1, device=self.dev)
        self.t_dev = torch.zeros(1, device=self.dev)
        self.tokens_d = torch.zeros(batch_size, seq_len, dtype=torch.int64, device=self.dev)
        self.prefix_d = torch.zeros(batch_size, self.eng.D, device=self.dev)
        self.prefix_n = torch.zeros(batch_size, self.eng.D, device=self.dev)
        self.stats = torch.zeros(4, device=self.dev)          # [n_valid, loss_sum, ., .] of the last step (global)
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
            # per-rank RNG stream for noise + dropout: same weights, different masks
            self.eng.seed.add_(7919 * (1 + torch.distributed.get_rank(process_group)))
        # data-parallel all-reduce overlapped with backward: one bucket per GPT-2 block, reduced on a side stream as
        # soon as the block's gradients are final; only [tail | mapper | wte | wpe] (ready last) stays exposed
        # (experimental, opt-in: NCCL collectives captured on a second graph branch were seen to dead-lock on replay)
        self.overlap = (self.world > 1 and self.train_gpt and os.environ.get("CAPDEC_DP_OVERLAP", "0") == "1")
        # CAPDEC_DP_OVERLAP=2 (experimental, opt-in, NOT yet validated on hardware): the same per-block buckets, but no
        # collective is ever captured.  The step's graph is CUT at every GPT-2 block boundary of the backward pass
        # (13 graphs: forward + head + block 11 | block 10 | ... | block 0 | embedding scatter + mapper) and the bucket of
        # the block a segment finished is all-reduced EAGERLY on the side stream between two segment replays, so every
        # rank issues its collectives from host code in one fixed order and the reduction of block l travels over NVLink
        # while blocks l-1 ... 0 are still computing.
        self.segmented = (self.world > 1 and self.train_gpt and os.environ.get("CAPDEC_DP_OVERLAP", "0") == "2")
        self._segs = None
        # opt-in (CAPDEC_DP_PIPELINE=1): the all-reduce is cut into chunks on a side stream and the fused AdamW of chunk k
        # runs while chunk k+1 is still being reduced (both outside the CUDA graph: plain stream/event ordering).
        # Validated on 2 GPUs (tests/test_dp_gpu.py passes with it) but measured equal to the plain sequence there
        # (19.00 vs 19.01 ms/step: the all-reduce and AdamW compete for the same HBM bandwidth), so it is not the default.
        self.pipeline = (self.world > 1 and not self.overlap and os.environ.get("CAPDEC_DP_PIPELINE", "0") == "1")
        # single GPU: the fused AdamW of a GPT-2 block runs on a second stream as soon as backward has finished that block
        # (its gradients are final and its weights are not read again this step), hidden behind the remaining backward
        # GEMMs - AdamW is pure HBM traffic (4.4 GB/step), the GEMMs are tensor / L2 bound.  [mapper | wte | wpe] follow at
        # the end.  Data parallel keeps reduce-then-update (the update needs the all-reduced gradient).
        # Opt-in (CAPDEC_OPT_OVERLAP=1): validated (tests/test_model_gpu.py etc. pass with it) but measured equal to the
        # plain sequence on this pool (17.79-17.89 vs 17.86 ms/step, gpurun_out/s15_bench_*.log) - the B200s here run
        # against their power cap, so overlapping HBM traffic with tensor work does not shorten the step.
        self.opt_overlap = (self.world == 1 and self.train_gpt and os.environ.get("CAPDEC_OPT_OVERLAP", "0") == "1")
        if self.opt_overlap:
            self.opt_stream = torch.cuda.Stream(device=self.dev)
        if self.train_gpt:     # parameter spans (in trainable-parameter coordinates) in the order backward finishes them
            lay = fl.layout
            nl = self.eng.nl
            first = lay["gpt.transformer.h.0.ln_1.weight"][0] - fl.tail
            self.layer_spans = []
            for l in range(nl):
                a0 = lay[f"gpt.transformer.h.{l}.ln_1.weight"][0] - fl.tail
                a1 = (lay[f"gpt.transformer.h.{l + 1}.ln_1.weight"][0] - fl.tail) if l + 1 < nl else self.n_train
                self.layer_spans.append((a0, a1))                 # block l (+ ln_f for the last block)
            self.head_span = (0, first)                           # mapper | wte | wpe: final only after the embedding scatter
        if self.overlap or self.pipeline or self.segmented:
            self.comm = torch.cuda.Stream(device=self.dev)
        if self.overlap or self.segmented:
            self.buckets, self.head_bucket = self.eng.layer_grad_slices()
            if not self.train_gpt or self.n_train != fl.grads.numel() - fl.tail:
                raise CapdecError("per-block gradient buckets need every parameter trainable")
        if self.pipeline:
            n_chunks = int(os.environ.get("CAPDEC_DP_CHUNKS", "4"))
            total = self.reduce_buf.numel()                     # 4 tail floats + trainable gradients
            step = ((total + n_chunks - 1) // n_chunks + 1023) // 1024 * 1024
            self.chunks = [(lo, min(total, lo + step)) for lo in range(0, total, step)]
        # Data parallel default: the optimizer is SHARDED over the ranks (ZeRO-1 style).  One reduce-scatter of the flat
        # gradient buffer gives every rank the summed gradients of its 1/N slice, the fused AdamW updates just that slice
        # (1/N of the 4.4 GB the update streams at C2), and one all-gather hands every rank the same new parameters - the
        # same NVLink traffic as the all-reduce it replaces (reduce-scatter + all-gather), the update N times cheaper,
        # Adam moments 1/N of the memory, and the replicas stay bit-identical by construction.  CAPDEC_DP_SHARDED=0 keeps
        # all-reduce + full update; so does a parameter count that does not split into N slices of whole float4s.
        self.sharded = (self.world > 1 and not (self.overlap or self.segmented or self.pipeline)
                        and os.environ.get("CAPDEC_DP_SHARDED", "1") != "0" and self.n_train % (4 * self.world) == 0)
        if self.sharded:
            rank = torch.distributed.get_rank(process_group)
            sh = self.n_train // self.world
            self.shard = (rank * sh, (rank + 1) * sh)
            self.m_flat, self.v_flat = torch.zeros(sh, device=self.dev), torch.zeros(sh, device=self.dev)
        # ... and by default the three steps are ONE kernel over NVLink peer memory (csrc/peer.cu): the owner of a slice
        # loads that slice of every rank's gradient buffer, updates, and stores the new parameters into every rank's
        # parameter buffer.  CAPDEC_DP_PEER=0 (or peers that cannot map each other's memory) keeps the NCCL sequence.
        # ... and the gradients reach their owner while the backward pass is still running: as soon as a GPT-2 block's
        # gradients are final, every rank's COPY ENGINES push its share of that bucket into the owners' staging areas
        # (memcpy nodes inside the step's CUDA graph).  No SM is involved - unlike an NCCL kernel, the transfer takes
        # nothing from the persistent GEMM CTAs - and the update kernel then reads all N contributions from local HBM:
        # the only exposed NVLink traffic of a step is the parameter all-gather.  CAPDEC_DP_PUSH=0: the update kernel
        # loads the gradients over NVLink itself (measured 365 GB/s per direction at N = 2: loads are latency-bound).
        self.peer = self.push = False
        if self.sharded and os.environ.get("CAPDEC_DP_PEER", "1") != "0" and self.world <= 8:
            want_push = (os.environ.get("CAPDEC_DP_PUSH", "1") != "0" and self.train_gpt
                         and self.n_train == fl.grads.numel() - fl.tail)
            with torch.cuda.device(self.dev):      # the handle exchange is a collective on THIS rank's device
                self.peer = self._peer_setup(want_push)
        self.use_graph = use_cuda_graph
        self._g_fb = self._g_opt = self._g_eval = None
        self._warm = 0
        self._eval_warm = 0
        self.eval_stats = torch.zeros(4, device=self.dev)
        fl.grads.zero_()

    # ---- pieces ------------------------------------------------------------------------------------------------
    def _fwd_bwd(self, layer_hook=None):
        eng = self.eng
        ops.step_clock(eng.seed, self.step_dev, self.lr_dev, self.t_dev, self.lr, self.warmup, self.total)
        if self.noise_variance > 0.0:
            ops.noise_injection(self.prefix_d, self.prefix_n, self.noise_variance, offset=self.offset,
                                uniform_ball=self.uniform_noise, dont_norm=self.dont_norm, seed=eng.seed)
            pfx = self.prefix_n
        else:
            pfx = self.prefix_d                                 # train.py:28-29: variance 0 -> untouched
        hook = self._reduce_layer if (self.overlap or self.segmented) else (self._opt_layer if self.opt_overlap else None)
        if self.push:
            hook = self._push_layer
        if layer_hook is not None:       # segmented capture: the hook cuts the graph instead of launching a collective
            hook = layer_hook
        self._stats_taken = False
        before_bwd = None
        if self.peer and layer_hook is None:
            # the gradient buffer is cleared on a side stream WHILE the forward pass runs (a memset node of the step graph)
            # instead of after the update: nobody reads it between the closing fence of the previous step and this point
            ev = torch.cuda.Event()
            ev.record()
            self.zero_stream.wait_event(ev)
            with torch.cuda.stream(self.zero_stream):
                ops.zero_fill(self.g_flat)
                zeroed = torch.cuda.Event()
                zeroed.record()
            before_bwd = lambda: torch.cuda.current_stream().wait_event(zeroed)
        eng.loss_and_grads(self.tokens_d, pfx, train_gpt=self.train_gpt, mean_reduce=False, on_layer_done=hook,
                           before_backward=before_bwd)
        if layer_hook is not None:
            return
        if self.push:          # [mapper | wte | wpe] became final last; then the step (graph) ends with every push issued
            self._push_span(self.head_span)
            torch.cuda.current_stream().wait_stream(self.push_stream)
        if self.opt_overlap:
            self._take_stats()
            self._adamw_span(*self.head_span)
            torch.cuda.current_stream().wait_stream(self.opt_stream)
        if self.overlap or self.segmented:
            # [tail | mapper | wte | wpe] is ready last.  It goes to the SAME side stream as the per-block buckets: every
            # collective of this communicator then sits on one branch of the captured graph, in one fixed order on every
            # rank (collectives on two concurrent branches may replay in different orders on different ranks and deadlock)
            self._reduce_bucket(self.head_bucket)
            torch.cuda.current_stream().wait_stream(self.comm)

    def _reduce_layer(self, l: int):
        self._reduce_bucket(self.buckets[l])

    def _reduce_bucket(self, bucket: torch.Tensor):
        ev = torch.cuda.Event()
        ev.record()
        self.comm.wait_event(ev)
        with torch.cuda.stream(self.comm):
            torch.distributed.all_reduce(bucket, group=self.pg)

    def _take_stats(self):
        if not self._stats_taken:      # [n_valid, loss_sum] are final once the CE kernel ran, i.e. before backward starts
            self.stats.copy_(self.tail)
            self._stats_taken = True

    def _adamw_span(self, lo: int, hi: int):
        if hi > lo:
            ops.adamw_step(self.p_flat[lo:hi], self.g_flat[lo:hi], self.m_flat[lo:hi], self.v_flat[lo:hi], self.lr_dev,
                           self.t_dev, self.betas[0], self.betas[1], self.eps, self.wd, grad_denom=self.stats[0:1],
                           zero_grad=True)

    def _opt_layer(self, l: int):
        self._take_stats()
        ev = torch.cuda.Event()
        ev.record()
        self.opt_stream.wait_event(ev)
        with torch.cuda.stream(self.opt_stream):
            self._adamw_span(*self.layer_spans[l])

    def _reduce_scatter_opt(self):
        """Sharded data-parallel update (see __init__): global counts, reduce-scatter, AdamW on this rank's slice,
        all-gather of the new parameters; the rest of the gradient buffer is cleared for the next step."""
        lo, hi = self.shard
        self.stats.copy_(self.tail)
        torch.distributed.all_reduce(self.stats, group=self.pg)                    # [n_valid, loss_sum, ., .] of all ranks
        torch.distributed.reduce_scatter_tensor(self.g_flat[lo:hi], self.g_flat, group=self.pg)   # in place (NCCL)
        ops.adamw_step(self.p_flat[lo:hi], self.g_flat[lo:hi], self.m_flat, self.v_flat, self.lr_dev, self.t_dev,
                       self.betas[0], self.betas[1], self.eps, self.wd, grad_denom=self.stats[0:1], zero_grad=True)
        torch.distributed.all_gather_into_tensor(self.p_flat, self.p_flat[lo:hi], group=self.pg)  # in place (NCCL)
        if lo > 0:
            ops.zero_fill(self.g_flat[:lo])
        if hi < self.n_train:
            ops.zero_fill(self.g_flat[hi:])

    def _peer_setup(self, want_push: bool) -> bool:
        """Exchange CUDA-IPC handles of the flat gradient / parameter buffers (and of the gradient staging area) and map
        every peer's (one process per GPU, one node).  Collective: every rank learns whether ALL ranks succeeded, so that
        all take the same path."""
        dist = torch.distributed
        rank = dist.get_rank(self.pg)
        sh = self.shard[1] - self.shard[0]
        staging = torch.empty((self.world - 1) * sh, device=self.dev) if want_push else None
        try:
            mine = (ops.peer_export(self.g_flat), ops.peer_export(self.p_flat),
                    ops.peer_export(staging) if want_push else None)
        except CapdecError as e:
            mine = str(e)
        table = [None] * self.world
        dist.all_gather_object(table, mine, group=self.pg)
        g_ptrs, p_ptrs, s_ptrs, opened, err = [], [], [], [], None
        if any(isinstance(x, str) for x in table):
            err = next(x for x in table if isinstance(x, str))
        else:
            try:
                for r, (gx, px, sx) in enumerate(table):
                    if r == rank:
                        g_ptrs.append(self.g_flat.data_ptr()); p_ptrs.append(self.p_flat.data_ptr())
                        s_ptrs.append(staging.data_ptr() if want_push else 0)
                        continue
                    for exp, dst in ((gx, g_ptrs), (px, p_ptrs), (sx, s_ptrs)):
                        if exp is None:
                            dst.append(0)
                            continue
                        ptr = ops.peer_open(*exp)
                        opened.append((ptr, exp[1]))
                        dst.append(ptr)
            except CapdecError as e:
                err = str(e)
        oks = [None] * self.world
        dist.all_gather_object(oks, err, group=self.pg)
        bad = [x for x in oks if x is not None]
        if bad:
            for ptr, off in opened:
                ops.peer_close(ptr, off)
            if rank == 0:
                print(f"capdec_b200: peer-memory optimizer step unavailable ({bad[0]}); using NCCL reduce-scatter / all-gather",
                      flush=True)
            return False
        self.g_ptrs, self.p_ptrs, self._peer_opened, self.rank = g_ptrs, p_ptrs, opened, rank
        self.fence = torch.zeros(4, device=self.dev)
        self.zero_stream = torch.cuda.Stream(device=self.dev)
        lo = self.shard[0]
        if want_push:
            self.push = True
            self.staging, self.staging_ptrs = staging, s_ptrs
            self.push_stream = torch.cuda.Stream(device=self.dev)
            # the update kernel reads rank r's contribution to this rank's slice from its local staging slot (own: in place)
            self.g_slices = [self.g_flat.data_ptr() + 4 * lo if r == rank else
                             staging.data_ptr() + 4 * sh * (r if r < rank else r - 1) for r in range(self.world)]
        else:
            self.g_slices = [ptr + 4 * lo for ptr in g_ptrs]      # straight from the peers' gradient buffers (NVLink loads)
        return True

    def _push_span(self, span):
        """Copy-engine pushes of this rank's gradients [span) into the staging areas of the ranks that own them, on the
        side stream, ordered after everything the main stream has issued so far (memcpy nodes under graph capture)."""
        plan = push_plan(span, self.world, self.rank, self.shard[1] - self.shard[0])
        if not plan:
            return
        ev = torch.cuda.Event()
        ev.record()
        self.push_stream.wait_event(ev)
        base = self.g_flat.data_ptr()
        with torch.cuda.stream(self.push_stream):
            for r, first, count, dst_off in plan:
                ops.copy_async(self.staging_ptrs[r] + 4 * dst_off, base + 4 * first, 4 * count)

    def _push_layer(self, l: int):
        self._push_span(self.layer_spans[l])

    def _peer_opt(self):
        """Sharded update in one kernel over peer memory.  The all-reduce of the global counts completes on a rank only
        after every rank has entered it, i.e. after every rank's backward pass (stream order): all gradients are final when
        the kernel starts.  The second 16-byte all-reduce is the matching fence at the other end: once it completes, every
        rank's kernel has finished - all new parameters have landed here, and nobody reads this rank's gradients any more."""
        lo, hi = self.shard
        self.stats.copy_(self.tail)
        torch.distributed.all_reduce(self.stats, group=self.pg)
        ops.adamw_peer_step(self.g_slices, self.p_ptrs, self.rank, lo, hi - lo, self.m_flat, self.v_flat, self.lr_dev,
                            self.t_dev, self.betas[0], self.betas[1], self.eps, self.wd, grad_denom=self.stats[0:1])
        torch.distributed.all_reduce(self.fence, group=self.pg)      # (the next step clears the gradients during its forward)

    def _opt(self):
        if self.opt_overlap:           # the update already ran inside _fwd_bwd
            return
        self.stats.copy_(self.tail)
        ops.adamw_step(self.p_flat, self.g_flat, self.m_flat, self.v_flat, self.lr_dev, self.t_dev, self.betas[0],
                       self.betas[1], self.eps, self.wd, grad_denom=self.stats[0:1], zero_grad=True)

    def autotune(self):
        """Measure the GEMM plans (engine / tile width / split-K) of this step's problems on the batch resident in
        `tokens_d` / `prefix_d`: one throw-away forward+backward with capdec_gemm_autotune on.  The measuring launches
        repeat every GEMM, so the gradients of that pass are garbage: they are discarded, and the step clock / RNG seed
        are restored.  Parameters and optimizer state are never touched.  CAPDEC_GEMM_AUTOTUNE=0 disables it."""
        if os.environ.get("CAPDEC_GEMM_AUTOTUNE", "1") == "0" or not ops.is_tc():
            return 0
        keep = [t.clone() for t in (self.eng.seed, self.step_dev, self.lr_dev, self.t_dev)]
        ops.gemm_autotune(1)
        self.eng.serial_backward = True     # one stream while measuring: concurrent kernels would distort the timings
        try:   # the same launches as _fwd_bwd, but never a collective: ranks measure independently
            pfx = self.prefix_d
            if self.noise_variance > 0.0:
                ops.noise_injection(self.prefix_d, self.prefix_n, self.noise_variance, offset=self.offset,
                                    uniform_ball=self.uniform_noise, dont_norm=self.dont_norm, seed=self.eng.seed)
                pfx = self.prefix_n
            self.eng.loss_and_grads(self.tokens_d, pfx, train_gpt=self.train_gpt, mean_reduce=False)
        finally:
            self.eng.serial_backward = False
            n = ops.gemm_autotune(0)
        torch.cuda.synchronize()
        for dst, src in zip((self.eng.seed, self.step_dev, self.lr_dev, self.t_dev), keep):
            dst.copy_(src)
        self.eng.flat.grads.zero_()
        return n

    def _reduce_and_opt(self):
        """Chunked all-reduce (side stream) pipelined with the fused AdamW (main stream).  Chunk 0 carries the tail
        [n_valid, loss_sum, ., .], so the global token count is known before the first parameter is updated."""
        main = torch.cuda.current_stream()
        tail = self.eng.flat.tail
        ready = torch.cuda.Event()
        ready.record(main)
        self.comm.wait_event(ready)
        for k, (lo, hi) in enumerate(self.chunks):
            with torch.cuda.stream(self.comm):
                torch.distributed.all_reduce(self.reduce_buf[lo:hi], group=self.pg)
                done = torch.cuda.Event()
                done.record(self.comm)
            main.wait_event(done)
            if k == 0:
                self.stats.copy_(self.tail)
            plo, phi = max(lo, tail) - tail, hi - tail           # the same span in parameter coordinates
            if phi > plo:
                ops.adamw_step(self.p_flat[plo:phi], self.g_flat[plo:phi], self.m_flat[plo:phi], self.v_flat[plo:phi],
                               self.lr_dev, self.t_dev, self.betas[0], self.betas[1], self.eps, self.wd,
                               grad_denom=self.stats[0:1], zero_grad=True)
        self.comm.wait_stream(main)   # the next step's all-reduce must not start before these updates were issued

    def _capture(self, fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def _capture_segments(self):
        """CAPDEC_DP_OVERLAP=2: capture forward+backward as a chain of graphs cut at the block boundaries of backward.
        `Engine._trunk_bwd` calls the hook once every gradient of block l is final and its side-stream work has been
        joined, which is exactly where one capture may end and the next begin.  No collective is captured."""
        segs = []
        pool = torch.cuda.graph_pool_handle()
        cap = torch.cuda.Stream(device=self.dev)
        cap.wait_stream(torch.cuda.current_stream())
        cur = {}

        def begin():
            cur["g"] = torch.cuda.CUDAGraph()
            cur["g"].capture_begin(pool=pool)

        def cut(l: int):
            cur["g"].capture_end()
            segs.append((cur["g"], self.buckets[l], self.layer_spans[l]))
            begin()

        with torch.cuda.stream(cap):
            begin()
            self._fwd_bwd(layer_hook=cut)
            cur["g"].capture_end()
            segs.append((cur["g"], self.head_bucket, self.head_span))
        torch.cuda.current_stream().wait_stream(cap)
        return segs

    def _replay_segments(self):
        """Segment replays on the main stream; on the side stream, per finished segment: all-reduce of that block's
        gradient bucket, then the fused AdamW of the same parameters - both run while the main stream is already replaying
        the backward of the blocks below.  The global token count the update divides by is all-reduced first (the CE
        kernel finishes before the first cut), so no update waits for the head bucket.  Exposed at the end of the step:
        the all-reduce + update of [mapper | wte | wpe], which the embedding scatter completes last."""
        main = torch.cuda.current_stream()
        self.comm.wait_stream(main)          # the previous step's forward has read the parameters the updates will overwrite
        for i, (g, bucket, (lo, hi)) in enumerate(self._segs):
            g.replay()
            ev = torch.cuda.Event()
            ev.record(main)
            self.comm.wait_event(ev)
            with torch.cuda.stream(self.comm):
                if i == 0:               # [n_valid, loss_sum] of this rank are final -> global counts
                    self.stats.copy_(self.tail)
                    torch.distributed.all_reduce(self.stats, group=self.pg)
                torch.distributed.all_reduce(bucket, group=self.pg)
                self._adamw_span(lo, hi)
        main.wait_stream(self.comm)

    def step_device(self):
        """One step on the batch already resident in `tokens_d` / `prefix_d`."""
        if self.use_graph and self._warm >= 2:
            if self._g_fb is None and self._segs is None:
                torch.cuda.synchronize()
                if self.segmented:
                    self._segs = self._capture_segments()
                else:
                    self._g_fb = self._capture(self._fwd_bwd)
                self._g_opt = None if (self.opt_overlap or self.segmented or self.sharded) else self._capture(self._opt)
            if self.segmented:
                self._replay_segments()      # reduces AND updates bucket by bucket
                return self.stats
            self._g_fb.replay()
            if self.peer:
                self._peer_opt()
            elif self.sharded:
                self._reduce_scatter_opt()
            elif self.pipeline:
                self._reduce_and_opt()
            elif not self.opt_overlap:
                if self.world > 1 and not self.overlap:
                    torch.distributed.all_reduce(self.reduce_buf, group=self.pg)
                self._g_opt.replay()
        else:
            self._warm += 1
            self._fwd_bwd()
            if self.peer:
                self._peer_opt()
            elif self.sharded:
                self._reduce_scatter_opt()
            elif self.pipeline:
                self._reduce_and_opt()
            else:
                if self.world > 1 and not (self.overlap or self.segmented):
                    torch.distributed.all_reduce(self.reduce_buf, group=self.pg)
                self._opt()
            if self._warm == 2:   # GEMM tile planning for the live rows seen during warm-up (one host sync, results unaffected)
                self.eng.measure_row_hints(self.B, self.L)
                self.autotune()
        return self.stats

    def step(self, tokens: torch.Tensor, prefix: torch.Tensor):
        """tokens int64 [B, L], prefix fp32 [B, D] (host pinned or device).  Returns the device stats tensor
        [n_valid, loss_sum, ., .]; `Trainer.loss()` turns it into the mean token loss (forces a sync)."""
        if tuple(tokens.shape) != (self.B, self.L) or prefix.shape[0] != self.B:
            raise CapdecError(f"Trainer was built for batch {self.B} x {self.L}, got tokens {tuple(tokens.shape)}")
        self.tokens_d.copy_(tokens, non_blocking=True)
        self.prefix_d.copy_(prefix, non_blocking=True)
        return self.step_device()

    # ---- validation pass (train.py:372-389) ------------------------------------------------------------------------
    def evaluate_device(self):
        """Forward + masked CE on the batch resident in `tokens_d` / `prefix_d` with the model in eval mode: no noise
        injection (train.py:382-383 feeds the raw prefix), no dropout, no gradients.  Returns the device tensor
        [n_valid, loss_sum, ., .] (summed over ranks under data parallel)."""
        was_training = self.model.training
        self.model.eval()
        try:
            if self.use_graph and self._eval_warm >= 1:
                if self._g_eval is None:
                    torch.cuda.synchronize()
                    self._g_eval = self._capture(lambda: self.eng.loss_only(self.tokens_d, self.prefix_d, self.eval_stats))
                self._g_eval.replay()
            else:
                self._eval_warm += 1
                self.eng.loss_only(self.tokens_d, self.prefix_d, self.eval_stats)
        finally:
            self.model.train(was_training)
        if self.world > 1:
            torch.distributed.all_reduce(self.eval_stats, group=self.pg)
        return self.eval_stats

    def evaluate(self, tokens: torch.Tensor, prefix: torch.Tensor) -> float:
        """Mean token loss of one validation batch, `nnf.cross_entropy(..., ignore_index=0)` of train.py:385."""
        if tuple(tokens.shape) != (self.B, self.L) or prefix.shape[0] != self.B:
            raise CapdecError(f"Trainer was built for batch {self.B} x {self.L}, got tokens {tuple(tokens.shape)}")
        self.tokens_d.copy_(tokens, non_blocking=True)
        self.prefix_d.copy_(prefix, non_blocking=True)
        s = self.evaluate_device().tolist()
        return s[1] / s[0] if s[0] > 0 else float("nan")

    # ---- device-resident data feed (capdec_b200.data.DeviceCaptionDataset) -----------------------------------------
    def step_from(self, dataset, idx: torch.Tensor):
        """One train step on the captions `idx` (int64 CUDA tensor [B]) of a DeviceCaptionDataset: the batch is gathered
        on the device (one launch) straight into the step's input buffers - no host work, no H2D copy."""
        dataset.gather(idx, self.tokens_d, self.prefix_d)
        return self.step_device()

    def evaluate_from(self, dataset, idx: torch.Tensor):
        dataset.gather(idx, self.tokens_d, self.prefix_d)
        return self.evaluate_device()

    def evaluate_any(self, dataset, idx: torch.Tensor):
        """Validation batch from a dataset whose max_seq_len (or batch size) differs from the training configuration
        (train.py:373-385 builds a separate ClipCocoDataset for --val_pt): same eval-mode forward + masked CE, run eagerly
        on buffers of that shape.  Falls through to the captured graph when the shapes match."""
        B, L = int(idx.numel()), int(dataset.max_seq_len)
        if (B, L) == (self.B, self.L):
            return self.evaluate_from(dataset, idx)
        bufs = self.__dict__.setdefault("_val_bufs", {})
        if (B, L) not in bufs:
            bufs.clear()                                   # one live validation shape: its activation arena is the big part
            self.eng.arenas = {k: v for k, v in self.eng.arenas.items() if (k[0], k[2]) == (self.B, self.L)}
            bufs[(B, L)] = (torch.zeros(B, L, dtype=torch.int64, device=self.dev), torch.zeros(B, self.eng.D, device=self.dev))
        tokens, prefix = bufs[(B, L)]
        dataset.gather(idx, tokens, prefix)
        was_training = self.model.training
        self.model.eval()
        try:
            self.eng.loss_only(tokens, prefix, self.eval_stats)
        finally:
            self.model.train(was_training)
        if self.world > 1:
            torch.distributed.all_reduce(self.eval_stats, group=self.pg)
        return self.eval_stats

    def loss(self) -> float:
        s = self.stats.tolist()
        return s[1] / s[0] if s[0] > 0 else float("nan")

    def loss_lagged(self) -> Optional[float]:
        """Per-step loss logging without stalling the device: enqueues the D2H copy of THIS step's [n_valid, loss_sum] into
        pinned memory and returns the mean token loss of the PREVIOUS step (None after the first step), waiting only for
        that earlier copy.  The host stays one step ahead of the GPU, so graph launches and collectives of step i+1 are
        issued while step i still computes; `loss()` after the last step reads the final value."""
        if not hasattr(self, "_lag"):
            self._lag = {"buf": [torch.zeros(4, pin_memory=True) for _ in range(2)],
                         "ev": [torch.cuda.Event() for _ in range(2)], "n": 0}
        lag = self._lag
        cur = lag["n"] & 1
        lag["buf"][cur].copy_(self.stats, non_blocking=True)
        lag["ev"][cur].record()
        lag["n"] += 1
        if lag["n"] == 1:
            return None
        lag["ev"][cur ^ 1].synchronize()
        s = lag["buf"][cur ^ 1].tolist()
        return s[1] / s[0] if s[0] > 0 else float("nan")
