"""The G+D training step around the CUDA-backed modules: the reference's call sequence, flat gradient /
parameter storage, a fused Adam(W) step, bucketed data-parallel all-reduce, and whole-step CUDA-graph capture.

Call sites mirrored: src/v2/training.py:177-211 and src/v1/gan.py:222-252 (identical 3 D passes + 1 G pass).
Everything here is host-side orchestration; the arithmetic lives in libvitgan_b200 (and torch's loss heads).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import functional as Fn
from . import ops


# --------------------------------------------------------------------------------------------------
# flat storage: parameters and gradients of one network as views into two flat fp32 buffers
# --------------------------------------------------------------------------------------------------
class FlatNet:
    """Re-homes the parameters (and .grad) of `module` into flat fp32 buffers WITHOUT replacing the
    nn.Parameter objects (names, shapes, optimizer references and state_dict stay intact).
    `exclude(name, param)` leaves a parameter out (e.g. the reference-frozen q/k/v of the v1 discriminator)."""

    def __init__(self, module, exclude=None):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad and not (exclude and exclude(n, p))]
        # matrices first, vectors after (stable): q/k/v weights -- and, separately, their biases -- of one attention
        # block then sit back to back, so the fused QKV weight gradient is ONE split-K GEMM into one flat-buffer view
        named = [x for x in named if x[1].dim() >= 2] + [x for x in named if x[1].dim() < 2]
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        # 16-byte aligned slots so every view is vector-load friendly
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.param_ids = {id(p) for p in self.params}
        self.flat_param = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat_param[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
        # bf16 shadow of the whole flat buffer (operands of the tensor-core GEMMs); FusedAdam keeps it in sync
        self.flat_shadow = None
        self._versions = None
        if dev.type == "cuda":
            self.flat_shadow = torch.empty(off, dtype=torch.bfloat16, device=dev)
            self.sync_shadow()
            Fn.register_flat(self, self.params, self.offsets)

    def sync_shadow(self):
        """Re-derive the bf16 shadow from the fp32 parameters (after construction / load_state_dict)."""
        if self.flat_shadow is not None:
            with torch.no_grad():
                ops.cast(self.flat_param, torch.bfloat16, out=self.flat_shadow)
            self._versions = [p._version for p in self.params]
            self._vmap = {id(p): i for i, p in enumerate(self.params)}

    def shadow_valid(self, params):
        """False if any of `params` was modified through autograd-visible in-place ops since the last sync
        (e.g. load_state_dict, a torch optimizer): callers then fall back to the cast path."""
        if self._versions is None:
            return True
        return all(self._versions[self._vmap[id(p)]] == p._version for p in params)

    def zero_grad(self):
        """Keeps .grad as views of the flat buffer (autograd then accumulates in place)."""
        self.flat_grad.zero_()
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)


class FusedAdam:
    """torch.optim.Adam / AdamW semantics (single launch over the flat buffers; vg_adam_step).
    Mirrors AdamW(lr, weight_decay=1e-3) of src/v2/training.py:150-157 and Adam(lr, betas=(0.5,0.999)) of
    src/v1/gan.py:316-328.  The step counter lives on the device so the update is graph-capturable."""

    def __init__(self, net: FlatNet, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        self.net, self.lr, self.betas, self.eps, self.wd, self.decoupled = net, lr, betas, eps, weight_decay, decoupled
        dev = net.flat_param.device
        self.exp_avg = torch.zeros_like(net.flat_param)
        self.exp_avg_sq = torch.zeros_like(net.flat_param)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.grad_scale = 1.0

    def zero_grad(self, set_to_none=False):
        self.net.zero_grad()

    def step(self):
        with torch.no_grad():
            ops.adam_step(self.net.flat_param, self.net.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr,
                          self.betas[0], self.betas[1], self.eps, self.wd, self.decoupled, self.grad_scale,
                          shadow=self.net.flat_shadow)
        Fn.invalidate_operands(self.net.param_ids)   # the kernel wrote parameter memory behind autograd's back


# --------------------------------------------------------------------------------------------------
# data parallel: bucketed gradient all-reduce, launched from backward hooks, overlapped with backward
# --------------------------------------------------------------------------------------------------
class GradBuckets:
    """Splits a FlatNet's gradient buffer into contiguous buckets (reverse registration order ~ backward
    order) and all-reduces a bucket asynchronously as soon as every gradient in it has been accumulated in
    an *armed* backward pass.  Works with NCCL (GPU, one process per GPU) and gloo (CPU tests).

    The D network is armed only for its second backward (grads of pass 1 accumulate locally);
    the G network for its only backward.  `finish()` waits and rescales by 1/world_size."""

    def __init__(self, net: FlatNet, n_buckets=2, group=None, average_in_place=True):
        self.net, self.group, self.average_in_place = net, group, average_in_place
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n_buckets = max(1, min(n_buckets, len(net.params)))
        per = (net.numel + n_buckets - 1) // n_buckets
        self.bucket_of, self.bounds = [], []
        # bucket boundaries on parameter boundaries
        start, b = 0, 0
        for i, (p, o) in enumerate(zip(net.params, net.offsets)):
            end = net.offsets[i + 1] if i + 1 < len(net.params) else net.numel
            self.bucket_of.append(b)
            if end - start >= per and i + 1 < len(net.params) and b + 1 < n_buckets:
                self.bounds.append((start, end)); start = end; b += 1
        self.bounds.append((start, net.numel))
        self.n = len(self.bounds)
        self.sizes = [0] * self.n
        for b in self.bucket_of:
            self.sizes[b] += 1
        self.armed = False
        self.pending = [0] * self.n
        self.works = []
        self.index = {id(p): i for i, p in enumerate(net.params)}
        self._hook_handles = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(net.params)]
        Fn.grad_ready_hooks.append(self._fused_ready)      # gradients accumulated in place by the kernels

    def close(self):
        """Detach from the parameters and from the fused-accumulation notification list (the hooks hold this object alive)."""
        for h in self._hook_handles:
            h.remove()
        self._hook_handles = []
        if self._fused_ready in Fn.grad_ready_hooks:
            Fn.grad_ready_hooks.remove(self._fused_ready)

    def _make_hook(self, i):
        def hook(param):
            if not self.armed or self.world == 1:
                return
            b = self.bucket_of[i]
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)
        return hook

    def _fused_ready(self, param):
        i = self.index.get(id(param))
        if i is not None:
            self._make_hook(i)(param)

    def _launch(self, b):
        Fn.join_param_grad_stream()      # side-stream weight-gradient GEMMs of this bucket must precede the all-reduce
        s, e = self.bounds[b]
        self.works.append(dist.all_reduce(self.net.flat_grad[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def arm(self):
        self.armed = True
        self.pending = list(self.sizes)
        self.works = []

    def finish(self):
        """Wait for the launched buckets, reduce any bucket whose hooks did not all fire (unused params), average."""
        if self.world > 1:
            for b in range(self.n):
                if self.pending[b] > 0:
                    self._launch(b)
            for w in self.works:
                w.wait()
            if self.average_in_place:       # else the optimizer folds 1/world into its update (FusedAdam.grad_scale)
                self.net.flat_grad.mul_(1.0 / self.world)
        self.armed = False
        self.works = []


# --------------------------------------------------------------------------------------------------
# the step
# --------------------------------------------------------------------------------------------------
def _push(name):
    """NVTX range around one phase of the step (the reference has no tracing; nsys / ncu --nvtx-include pick these up).
    A push/pop costs ~0.1 us without a profiler attached and is legal during graph capture."""
    torch.cuda.nvtx.range_push(name)


def _pop():
    torch.cuda.nvtx.range_pop()


def _bce(prob, target, reduction="mean"):
    """nn.BCELoss(reduction='mean') (src/v1/gan.py:16-20): on the GPU the one-launch fused head (vg_bce: loss + dprob),
    on CPU tensors (host-logic tests only) torch's."""
    if prob.is_cuda and reduction == "mean":
        return Fn.bce(prob, target)[0]
    return F.binary_cross_entropy(prob, target, reduction=reduction)


def gan_step(gen, disc, gen_opt, disc_opt, real, noise, loss_kind="ce", d_buckets=None, g_buckets=None,
             skip_unused_d_grads=False, merge_d_passes=False):
    """One adversarial iteration: D(real), G(noise), D(fake.detach()) -> D step; D(fake) -> G step.

    loss_kind 'ce'  : nn.CrossEntropyLoss with class-index targets (v2; shim Q2)
    loss_kind 'bce' : nn.BCELoss with float (B,1) targets (v1)
    skip_unused_d_grads: do not produce D's parameter gradients in the third pass (the reference computes
    and then discards them, training.py:177 / gan.py:222); dgrad still flows to G.  Off by default.
    merge_d_passes: run the two discriminator-update passes D(real), D(fake.detach()) as ONE pass over the
    concatenated batch with loss = mean_real + mean_fake.  No op on the path couples samples (no BatchNorm), so the
    accumulated gradient is the same sum up to fp32 summation order; every kernel sees 2x the rows, which halves the
    per-launch fixed cost of these passes (the E=128 configs are launch/latency bound).  Off by default.
    """
    b, dev = real.shape[0], real.device
    if loss_kind == "ce":
        ones = torch.ones(b, dtype=torch.long, device=dev)
        zeros = torch.zeros(b, dtype=torch.long, device=dev)
        crit = F.cross_entropy
    else:
        ones = torch.ones(b, 1, device=dev)
        zeros = torch.zeros(b, 1, device=dev)
        crit = _bce
    fused_ce = loss_kind == "ce" and merge_d_passes      # the optimised step also uses the one-launch CE head (vg_softmax_ce)
    _push("vg.d_update")
    disc_opt.zero_grad(set_to_none=False) if isinstance(disc_opt, FusedAdam) else disc_opt.zero_grad(set_to_none=True)
    if merge_d_passes:
        fake = gen(noise)
        if d_buckets is not None:
            d_buckets.arm()
        out = disc(torch.cat([real, fake.detach().to(real.dtype)], 0)).float()
        if fused_ce and out.is_cuda:
            per = Fn.softmax_ce(out, torch.cat([ones, zeros], 0), b)          # [loss_real, loss_fake], one launch
            loss_real, loss_fake = per[0], per[1]
            per.sum().backward()
        elif loss_kind == "bce" and out.is_cuda:
            per = Fn.bce(out, torch.cat([ones, zeros], 0), b)                 # fused BCELoss head (vg_bce), one launch
            loss_real, loss_fake = per[0], per[1]
            per.sum().backward()
        else:
            per = crit(out, torch.cat([ones, zeros], 0), reduction="none").view(2, -1)
            loss_real, loss_fake = per[0].mean(), per[1].mean()
            (loss_real + loss_fake).backward()
    else:
        loss_real = crit(disc(real).float(), ones)
        loss_real.backward()
        fake = gen(noise)
        if d_buckets is not None:
            d_buckets.arm()
        loss_fake = crit(disc(fake.detach()).float(), zeros)
        loss_fake.backward()
    if d_buckets is not None:
        d_buckets.finish()
    Fn.join_param_grad_stream()
    _push("vg.d_optimizer")
    disc_opt.step()
    _pop()
    _pop()
    _push("vg.g_update")
    gen_opt.zero_grad(set_to_none=False) if isinstance(gen_opt, FusedAdam) else gen_opt.zero_grad(set_to_none=True)
    if g_buckets is not None:
        g_buckets.arm()
    with Fn.skip_param_grads(skip_unused_d_grads):
        out = disc(fake)
    if fused_ce and out.is_cuda:
        loss_g = Fn.softmax_ce(out.float(), ones, b)[0]
    else:
        loss_g = crit(out.float(), ones)
    loss_g.backward()     # with skip_unused_d_grads only D's blocks were recorded with skip_pg; G's are unaffected
    if g_buckets is not None:
        g_buckets.finish()
    Fn.join_param_grad_stream()
    _push("vg.g_optimizer")
    gen_opt.step()
    _pop()
    _pop()
    return loss_real.detach(), loss_fake.detach(), loss_g.detach()


def plan_keep_g_graphs(gen, disc, real_mb, noise_mb, n_micro, reserve_bytes=12 << 30, info=None):
    """How many micro-batches' generator graphs `gan_step_microbatched(keep_g_graphs=...)` can keep alive across the discriminator
    update: measured, not guessed -- one eager generator forward (with grad) gives the bytes one kept graph pins, one
    discriminator forward+backward on its output the transient peak beside it; the rest of the device's free memory
    (minus `reserve_bytes` for the optimizer kernels, NCCL and allocator slack) is divided by the former.  `info`, if given, receives
    the measured sizes."""
    if n_micro <= 1:
        return 0
    dev = real_mb.device
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
    base = torch.cuda.memory_allocated(dev)
    fake = gen(noise_mb)
    torch.cuda.synchronize(dev)
    g_bytes = torch.cuda.memory_allocated(dev) - base
    torch.cuda.reset_peak_memory_stats(dev)
    disc(fake.detach()).float().sum().backward()
    Fn.join_param_grad_stream()
    torch.cuda.synchronize(dev)
    d_peak = torch.cuda.max_memory_allocated(dev) - base - g_bytes
    del fake
    for p in disc.parameters():      # the probe's gradients are not part of any step
        if p.grad is not None:
            p.grad.zero_()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info(dev)
    room = free - int(1.25 * d_peak) - g_bytes - reserve_bytes      # one graph is alive in the generator update anyway
    k = int(max(0, min(n_micro, room // max(g_bytes, 1))))
    if info is not None:
        info.update(graph_gb=round(g_bytes / 2 ** 30, 2), d_pass_peak_gb=round(d_peak / 2 ** 30, 2), free_gb=round(free / 2 ** 30, 2),
                    reserve_gb=round(reserve_bytes / 2 ** 30, 2), kept=k)
    return k


def gan_step_microbatched(gen, disc, gen_opt, disc_opt, real, noise, loss_kind="ce", n_micro=1, d_buckets=None, g_buckets=None,
                          skip_unused_d_grads=False, merge_d_passes=False, keep_g_graphs=0):
    """The same G+D iteration with the batch processed in `n_micro` equal chunks and exact mean-gradient accumulation
    (each chunk's loss is scaled by 1/n_micro; no BatchNorm on the path, so the result equals the full-batch step up to
    summation order).  Needed when the activations of the full batch do not fit (C4: global batch 2048 on one GPU).

    Phase D: per chunk  D(real) bwd, G(noise) without grad -> fake, D(fake) bwd;  then one D optimizer step.
    Phase G: per chunk  G(noise) with grad (G is unchanged, so this is the same fake), D(fake) bwd;  then one G step.
    Cost vs. the un-chunked step: one extra generator forward per chunk -- except for the first `keep_g_graphs` chunks, whose
    generator forward in phase D runs WITH grad and is kept (output + saved activations) until phase G reuses it: G does not
    change between the two phases, so the values are the ones the second forward would recompute; 180 GB of HBM hold the graphs
    of several 256-image chunks of the scaled config (`plan_keep_g_graphs` measures how many)."""
    if n_micro == 1:
        return gan_step(gen, disc, gen_opt, disc_opt, real, noise, loss_kind, d_buckets, g_buckets, skip_unused_d_grads, merge_d_passes)
    b, dev = real.shape[0], real.device
    assert b % n_micro == 0, "batch must be divisible by n_micro"
    mb = b // n_micro
    if loss_kind == "ce":
        ones, zeros, crit = torch.ones(mb, dtype=torch.long, device=dev), torch.zeros(mb, dtype=torch.long, device=dev), F.cross_entropy
    else:
        ones, zeros, crit = torch.ones(mb, 1, device=dev), torch.zeros(mb, 1, device=dev), _bce
    inv = 1.0 / n_micro
    zg = lambda opt: opt.zero_grad(set_to_none=False) if isinstance(opt, FusedAdam) else opt.zero_grad(set_to_none=True)
    _push("vg.d_update")
    zg(disc_opt)
    l_real = l_fake = l_g = 0.0
    kept = []
    for i in range(n_micro):
        r, z = real[i * mb:(i + 1) * mb], noise[i * mb:(i + 1) * mb]
        loss = crit(disc(r).float(), ones) * inv
        loss.backward()
        l_real = l_real + loss.detach()
        if i < keep_g_graphs:
            fake_g = gen(z)              # with grad: the graph is reused by the generator update below
            kept.append(fake_g)
            fake = fake_g.detach()
        else:
            with torch.no_grad():
                fake = gen(z)
        if d_buckets is not None and i == n_micro - 1:
            d_buckets.arm()
        loss = crit(disc(fake).float(), zeros) * inv
        loss.backward()
        l_fake = l_fake + loss.detach()
        if i + 1 < n_micro:
            Fn.join_param_grad_stream()     # releases the activations the side-stream weight-gradient GEMMs keep alive (one micro-batch at a time)
    if d_buckets is not None:
        d_buckets.finish()
    Fn.join_param_grad_stream()
    _push("vg.d_optimizer")
    disc_opt.step()
    _pop()
    _pop()
    _push("vg.g_update")
    zg(gen_opt)
    for i in range(n_micro):
        z = noise[i * mb:(i + 1) * mb]
        if g_buckets is not None and i == n_micro - 1:
            g_buckets.arm()
        if i < len(kept):
            fake, kept[i] = kept[i], None     # the kept graph; freed by its backward
        else:
            fake = gen(z)
        with Fn.skip_param_grads(skip_unused_d_grads):
            out = disc(fake)
        loss = crit(out.float(), ones) * inv
        loss.backward()
        l_g = l_g + loss.detach()
        fake = out = loss = None              # nothing of this chunk's graph outlives its backward
        if i + 1 < n_micro:
            Fn.join_param_grad_stream()
    if g_buckets is not None:
        g_buckets.finish()
    Fn.join_param_grad_stream()
    _push("vg.g_optimizer")
    gen_opt.step()
    _pop()
    _pop()
    return l_real, l_fake, l_g


class GraphedStep:
    """Whole-step CUDA graph: the ~600 kernel launches of one G+D iteration are captured once and replayed
    with one cudaGraphLaunch (the small-E configs are launch-bound).  Inputs are copied
    into static buffers; losses are read from static outputs.  Requires FusedAdam (device-side step counter)."""

    def __init__(self, gen, disc, gen_opt, disc_opt, real, noise, loss_kind="ce", warmup=3, param_grad_stream=True, **kw):
        self.real, self.noise = real.clone(), noise.clone()
        prev_pg = Fn.param_grad_stream_enabled()
        Fn.set_param_grad_stream(param_grad_stream)       # weight-gradient GEMMs on a forked stream inside the graph
        self.args = (gen, disc, gen_opt, disc_opt)
        self.kw = dict(loss_kind=loss_kind, **kw)
        step = gan_step_microbatched if kw.get("n_micro", 1) > 1 else gan_step
        if step is gan_step:
            self.kw.pop("n_micro", None)
        prev_cache = Fn.operand_cache_enabled()
        Fn.set_operand_cache(False)        # casts/packs must be part of the captured work
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                step(*self.args, self.real, self.noise, **self.kw)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()           # the warm-up's cached blocks belong to other streams / pools than the capture's private pool
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph):
                self.losses = step(*self.args, self.real, self.noise, **self.kw)
        finally:
            Fn.set_param_grad_stream(prev_pg)             # the fork is baked into the graph; eager callers keep their setting
            Fn.set_operand_cache(prev_cache)              # eager callers get their cached operand copies back
        # The graph holds raw device addresses (kernel arguments, tensor maps) of every parameter, gradient and optimizer buffer:
        # remember them, so that a replay after one of them was re-allocated (load onto another device, .to(dtype), a new
        # .grad tensor from zero_grad(set_to_none=True), ...) is an error instead of a silent write into freed memory.
        self._captured, self._captured_flat, self._calls = self._addresses(True), self._addresses(False), 0

    def _addresses(self, full):
        """Device addresses the graph depends on.  full: every parameter and gradient tensor (~1 us each: 0.2 ms for the 400
        tensors of the default GAN, so not on every replay of a 4 ms step); else only the flat buffers they are views of."""
        gen, disc, gen_opt, disc_opt = self.args
        addr = []
        if full:
            for net in (gen, disc):
                for p in net.parameters():
                    addr.append(p.data_ptr())
                    addr.append(p.grad.data_ptr() if p.grad is not None else 0)
        for opt in (gen_opt, disc_opt):
            if isinstance(opt, FusedAdam):
                addr.extend(t.data_ptr() for t in (opt.net.flat_param, opt.net.flat_grad, opt.exp_avg, opt.exp_avg_sq, opt.step_count))
        return addr

    def __call__(self, real, noise):
        # flat buffers on every replay, every tensor on the first replays and then every 64th
        full = self._calls < 2 or self._calls % 64 == 0
        self._calls += 1
        if self._addresses(False) != self._captured_flat or (full and self._addresses(True) != self._captured):
            raise RuntimeError("vitgan_b200.GraphedStep: a parameter, gradient or optimizer buffer was re-allocated after capture; "
                               "the captured graph still points at the old storage -- build a new GraphedStep")
        self.real.copy_(real, non_blocking=True)
        self.noise.copy_(noise, non_blocking=True)
        self.graph.replay()
        return self.losses
