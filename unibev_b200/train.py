"""Data-parallel training step of the hot path (BASELINE configs[4]: unibev_nus_LC_cat_128, batch sharded over the GPUs of
one box, gradient all-reduce over NCCL / NVLink).

The reference trains through mmcv's ``MMDistributedDataParallel`` (tools/train_UniBEV.py -> mmdet3d ``train_model``): one
process per GPU, ``samples_per_gpu`` samples each, gradients averaged over ranks before the optimizer step, every rank
seeded identically (``set_random_seed(args.seed)``, train_UniBEV.py:200-204) so the modality-dropout flags
(transformer_fusion.py:227-228, 474-477) coincide.  ``GradBuckets`` is that gradient exchange, sized for this path:

* parameters are grouped into buckets in REVERSE registration order (the order autograd finishes them: the last
  encoder layer first), one contiguous fp32 buffer per bucket;
* a bucket's all-reduce is launched on a side stream the moment its last gradient has been accumulated
  (``register_post_accumulate_grad_hook``), so the exchange of layer i overlaps the backward kernels of layer i - 1;
* buckets go to the collective strictly in bucket order (each once it and all earlier ones are complete), and
  parameters that received no gradient this step (a dropped modality's encoder) contribute zeros at ``finish()``, so
  every rank issues the same collectives in the same order whatever its flags;
* ``finish()`` joins the side stream, scales by 1 / world and scatters the averages back into ``p.grad``;
* ``prepare()`` (called by ``train_step`` before the backward pass) makes every ``p.grad`` a VIEW of its bucket slot, so
  autograd accumulates straight into the flat buffers and the all-reduced averages are the gradients: no per-parameter
  copy kernels in the hooks or in ``finish()`` (~2 x #parameters small launches per step otherwise, which is what an
  8-process box notices first: its host cores are shared by the ranks' launch threads);
* ``uniform_usage=True``: the ranks are seeded alike (as the reference's launcher does), so a parameter untouched here is
  untouched everywhere and ``finish()`` needs neither the "touched" flag exchange nor its host read-back.

The forward / backward arithmetic is the module path of ``unibev_b200.plugin`` (``ub_msda_fwd`` / ``ub_msda_bwd`` through
``ops.MultiScaleDeformableAttnFunction``).
"""
import torch
import torch.distributed as dist


class GradBuckets:
    def __init__(self, params, bucket_bytes=8 << 20, group=None, uniform_usage=False):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.uniform_usage = uniform_usage
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.buckets = []          # [(flat buffer, [(param, offset, numel)])]
        cur, cur_n = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_n += p.numel()
            if cur_n * 4 >= bucket_bytes:
                self._close(cur)
                cur, cur_n = [], 0
        if cur:
            self._close(cur)
        self._where = {}
        for bi, (_, items) in enumerate(self.buckets):
            for p, off, n in items:
                self._where[p] = (bi, off, n)
        self._pending = [len(items) for _, items in self.buckets]
        self._seen = set()
        self._next = 0             # next bucket to hand to the collective
        self._works = []
        self._stream = torch.cuda.Stream() if self.params and self.params[0].is_cuda else None
        self.defer_launch = False      # True: the hooks only book-keep, nothing goes on the wire before finish() / exchange_all()
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        # NCCL averages inside the collective; gloo has no AVG: sum, then one scale per bucket
        self._avg = (self.world > 1 and dist.get_backend(group) == 'nccl')

    def _view(self, p):
        bi, off, n = self._where[p]
        return self.buckets[bi][0][off:off + n].view_as(p)

    def _is_view(self, p):
        bi, off, n = self._where[p]
        flat = self.buckets[bi][0]
        return (p.grad is not None and p.grad.data_ptr() == flat.data_ptr() + off * 4 and p.grad.is_contiguous()
                and p.grad.dtype == torch.float32)

    def prepare(self):
        """Before a backward pass: zero the flat buffers and point every ``p.grad`` at its slot.  Optional -- a parameter
        whose ``.grad`` is anything else when its hook fires is copied into the slot as before."""
        if self.world == 1:
            return
        for flat, items in self.buckets:
            flat.zero_()
            for p, off, n in items:
                p.grad = flat[off:off + n].view_as(p)

    def _close(self, plist):
        n = sum(p.numel() for p in plist)
        flat = torch.zeros(n, dtype=torch.float32, device=plist[0].device)
        items, off = [], 0
        for p in plist:
            items.append((p, off, p.numel()))
            off += p.numel()
        self.buckets.append((flat, items))

    # ---- backward-time side ---------------------------------------------------------------------------------
    def _on_grad(self, p):
        """One backward pass per ``finish()``: a second gradient for a parameter whose bucket is already on the wire cannot
        be exchanged any more and raises; before that (gradient accumulation inside one backward graph, shared
        parameters) the accumulated ``p.grad`` is simply copied again."""
        if self.world == 1:
            return                                     # single process: gradients stay where autograd left them
        bi, off, n = self._where[p]
        view = self._is_view(p)
        if p in self._seen:
            if bi < self._next:
                raise RuntimeError('GradBuckets: a parameter received another gradient after its bucket was all-reduced; '
                                   'call finish() after every backward pass (no gradient accumulation across passes)')
            if not view:
                self.buckets[bi][0][off:off + n].copy_(p.grad.reshape(-1))
            return
        self._seen.add(p)
        flat = self.buckets[bi][0]
        if not view:
            flat[off:off + n].copy_(p.grad.reshape(-1))
        self._pending[bi] -= 1
        # collectives must be issued in the same order on every rank: bucket order, each as soon as it and all
        # earlier buckets are complete (a bucket of parameters this rank never touches waits for finish())
        while not self.defer_launch and self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._launch(self._next)
            self._next += 1

    def _launch(self, bi):
        flat = self.buckets[bi][0]
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                self._works.append(dist.all_reduce(flat, op=op, group=self.group, async_op=True))
        else:
            self._works.append(dist.all_reduce(flat, op=op, group=self.group, async_op=True))

    # ---- after backward -------------------------------------------------------------------------------------
    def finish(self):
        """Completes the exchange: buckets whose parameters got no gradient this step are reduced too (zeros), in bucket
        order, so all ranks run identical collectives; then p.grad <- average over ranks.  A parameter no rank used keeps
        ``p.grad = None`` (one small all-reduce of per-parameter "touched" flags decides), so weight decay / Adam leave it
        alone exactly as in a single-process run."""
        if self.world == 1:
            return
        for bi in range(self._next, len(self.buckets)):
            flat, items = self.buckets[bi]
            for p, off, n in items:
                if p not in self._seen and not self._is_view(p):      # (a view's slot was zeroed by prepare())
                    flat[off:off + n].zero_()
            self._launch(bi)
        self._next = 0
        if self.uniform_usage:
            used = {p: (1.0 if p in self._seen else 0.0) for p in self.params}
        else:
            dev = self.buckets[0][0].device
            touched = torch.tensor([1.0 if p in self._seen else 0.0 for p in self.params], device=dev)
            dist.all_reduce(touched, group=self.group)
        for w in self._works:
            w.wait()
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        if not self.uniform_usage:
            used = dict(zip(self.params, touched.tolist()))
        inv = 1.0 / self.world
        for flat, items in self.buckets:
            if not self._avg:
                flat.mul_(inv)
            for p, off, n in items:
                if used[p] == 0.0:
                    p.grad = None
                    continue
                if self._is_view(p):      # autograd accumulated into the slot: the average is already p.grad
                    continue
                g = flat[off:off + n].view_as(p)
                if p.grad is None:        # unused on this rank, used on another: the average over ranks
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
        self._works.clear()
        self._seen.clear()
        self._pending = [len(items) for _, items in self.buckets]

    def reset_pass(self):
        """Forget the current backward pass's book-keeping without exchanging anything (after a pass that was only recorded:
        ``GraphedTrainStep`` with ``exchange='after'``)."""
        self._works.clear()
        self._seen.clear()
        self._pending = [len(items) for _, items in self.buckets]
        self._next = 0

    def exchange_all(self):
        """All-reduce (average) every bucket in place, on the current stream, in bucket order.  For gradients that already
        live in the buckets (``prepare()``'s views) and a parameter set that is the same on every rank and step -- the
        replayed-graph step: no hooks ran, so there is no per-pass book-keeping to consult."""
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        for flat, _ in self.buckets:
            dist.all_reduce(flat, op=op, group=self.group)
            if not self._avg:
                flat.mul_(1.0 / self.world)

    def nbytes(self):
        return sum(flat.numel() * 4 for flat, _ in self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def train_step(model, bev_embedding, inputs, optimizer, buckets, loss_fn=None):
    """One step of the hot path: forward (module path, dropout + modality dropout active), scalar loss, backward with the
    gradient exchange overlapped, optimizer step.  ``bev_embedding`` is the learnable (Nq, C) query table
    (``bev_embedding.weight`` of the head, unibev_head.py:126-133).  Returns the detached loss."""
    optimizer.zero_grad(set_to_none=True)
    buckets.prepare()
    fused = model.encode(inputs['img_feats'], inputs['pts_feats'], bev_embedding, inputs['bev_h'], inputs['bev_w'],
                         bev_pos=inputs['bev_pos'], img_metas=inputs['img_metas'])
    loss = loss_fn(fused) if loss_fn is not None else fused.square().mean()
    loss.backward()
    buckets.finish()
    optimizer.step()
    return loss.detach()


class GraphedTrainStep:
    """``train_step`` replayed from CUDA graphs: the ~1 100 kernel launches of a step (forward, backward, gradient
    exchange, optimizer) become one ``cudaGraphLaunch``, so the step runs at the GPU's pace instead of the Python
    interpreter's (16.6 -> ~13 ms on one B200; with eight ranks sharing a host the gap is larger).

    * Inputs live in static device buffers (``img_feats``, ``pts_feats``, ``bev_pos``, ``lidar2img``); ``__call__`` copies
      the step's tensors into them (host tensors: pinned, ``non_blocking``) and replays.
    * The modality-dropout flags (transformer_fusion.py:227-228, 474-477) are drawn on the host exactly as the eager step
      draws them (same ``np.random`` stream); they only change two multipliers of the fusion, so there is one graph per
      flag pair, captured the first time the pair comes up (at most three).
    * ``exchange='in_graph'`` (one process): forward, backward and the optimizer step are one graph.  ``'after'`` (several
      ranks): the graph ends after backward -- gradients have been accumulated straight into the bucket buffers -- then
      the buckets are all-reduced in place (2-3 NCCL calls, ~25 MB: ~0.2 ms over NVLink, not overlapped) and the optimizer
      steps eagerly (one fused kernel).  ``'auto'`` picks by world size.  (Capturing the NCCL calls inside the graph hung
      on the 2-GPU box of this pod, so the exchange stays outside.)
    * The optimizer must be graph-capturable (``torch.optim.AdamW(..., fused=True, capturable=True)``).
    * Capturing a flag pair rehearses the step ``warmup`` times first (allocator pools, optimizer state, communicators);
      parameters and optimizer state are restored afterwards, so the sequence of updates is the eager one (dropout masks
      differ: the rehearsals advance the CUDA generator).

    Everything inside is the module path of the plugin: no tensor is created from host data and nothing synchronises while
    capturing (constant tensors come from ``ops.const_tensor`` / the encoders' grid cache, calibration from ``lidar2img``)."""

    def __init__(self, model, bev_embedding, optimizer, buckets, example, loss_fn=None, exchange='auto', warmup=3):
        if exchange not in ('auto', 'in_graph', 'after'):
            raise ValueError("exchange must be 'auto', 'in_graph' or 'after'")
        if exchange == 'auto':        # one process: nothing to exchange, the optimizer step joins the graph
            exchange = 'in_graph' if buckets.world == 1 else 'after'
        if exchange == 'in_graph' and buckets.world > 1:
            raise ValueError("exchange='in_graph' is for single-process runs; with several ranks the NCCL all-reduces run "
                             "after the graph (exchange='after')")
        self.model, self.emb, self.opt, self.buckets = model, bev_embedding, optimizer, buckets
        self.loss_fn, self.exchange, self.warmup = loss_fn, exchange, warmup
        dev = bev_embedding.device
        self.bev_h, self.bev_w = example['bev_h'], example['bev_w']
        self.static = {}
        for k in ('img_feats', 'pts_feats'):
            self.static[k] = [torch.empty_like(t, device=dev) for t in example[k]] if example.get(k) is not None else None
        self.static['bev_pos'] = torch.empty_like(example['bev_pos'], device=dev) if example.get('bev_pos') is not None else None
        self.static['lidar2img'], self.img_shape = None, None
        if example.get('img_metas') is not None:
            import numpy as np
            l2i = np.asarray([m['lidar2img'] for m in example['img_metas']], dtype=np.float32)
            self.static['lidar2img'] = torch.empty(l2i.shape, device=dev, dtype=torch.float32)
            self.img_shape = tuple(example['img_metas'][0]['img_shape'][0][:2])
        self._l2i_host = None
        self.graphs = {}
        self.captures = 0
        self.launches = {}            # flag pair -> libunibev_b200 kernel launches recorded in its graph
        self.replayed_launches = 0    # ... summed over the replays so far

    # ---- one step on the current stream with the static buffers ---------------------------------------------
    def _forward_backward(self, flags):
        fused = self.model.encode(self.static['img_feats'], self.static['pts_feats'], self.emb, self.bev_h, self.bev_w,
                                  bev_pos=self.static['bev_pos'], flags=flags, lidar2img=self.static['lidar2img'],
                                  img_shape=self.img_shape)
        loss = self.loss_fn(fused) if self.loss_fn is not None else fused.square().mean()
        loss.backward()
        return loss.detach()

    def _eager(self, flags):
        self.opt.zero_grad(set_to_none=True)
        self.buckets.prepare()
        loss = self._forward_backward(flags)
        self.buckets.finish()
        self.opt.step()
        return loss

    def _capture(self, flags):
        import numpy as np
        rng = np.random.get_state()              # encode() draws (and then overrides) the flags: keep the host stream as it was
        params = [p for group in self.opt.param_groups for p in group['params']]
        snap_p = [p.detach().clone() for p in params]
        snap_o = {p: {k: v.clone() for k, v in st.items() if torch.is_tensor(v)} for p, st in self.opt.state.items()}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):            # optimizer state, allocator pools, constant caches, NCCL communicators
            for _ in range(self.warmup):
                self._eager(flags)
            # the warm-up steps were rehearsals: parameters and optimizer state go back to where they were (in place: the
            # graph binds to these very tensors); state created by the rehearsal starts from zero like fresh state
            with torch.no_grad():
                for p, sp in zip(params, snap_p):
                    p.copy_(sp)
                for p, st in self.opt.state.items():
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            if p in snap_o and k in snap_o[p]:
                                v.copy_(snap_o[p][k])
                            else:
                                v.zero_()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        from . import _cabi
        g = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        n0 = _cabi.launch_count()
        self.buckets.defer_launch = self.exchange == 'after'
        try:
            with torch.cuda.graph(g):
                self.buckets.prepare()
                loss = self._forward_backward(flags)
                if self.exchange == 'in_graph':
                    self.buckets.finish()
                    self.opt.step()
        finally:
            self.buckets.defer_launch = False
        if self.exchange == 'after':
            self.buckets.reset_pass()
        self.launches[flags] = _cabi.launch_count() - n0
        np.random.set_state(rng)
        self.captures += 1
        self.graphs[flags] = (g, loss)
        return self.graphs[flags]

    def capture_all(self, inputs):
        """Captures every flag pair the model can draw (so that no capture lands inside a timed region)."""
        m = self.model
        has_img, has_pts = self.static['img_feats'] is not None, self.static['pts_feats'] is not None
        pairs = [(int(has_img), int(has_pts))]
        if m.drop_modality is not None and m.training and has_img and has_pts:
            pairs += [(1, 0), (0, 1)]
        self._stage(inputs)
        for flags in pairs:
            if flags not in self.graphs:
                self._capture(flags)

    def _stage(self, inputs):
        for k in ('img_feats', 'pts_feats'):
            if self.static[k] is not None:
                for dst, src in zip(self.static[k], inputs[k]):
                    dst.copy_(src, non_blocking=True)
        if self.static['bev_pos'] is not None:
            self.static['bev_pos'].copy_(inputs['bev_pos'], non_blocking=True)
        if self.static['lidar2img'] is not None:
            import numpy as np
            l2i = np.asarray([m['lidar2img'] for m in inputs['img_metas']], dtype=np.float32)
            if self._l2i_host is None:
                self._l2i_host = torch.empty(l2i.shape, dtype=torch.float32).pin_memory()
            self._l2i_host.copy_(torch.from_numpy(l2i))
            self.static['lidar2img'].copy_(self._l2i_host, non_blocking=True)

    def __call__(self, inputs):
        """One training step on ``inputs`` (same dict as ``train_step``); returns the detached loss (a static tensor that
        the next replay overwrites)."""
        m = self.model
        m._draw_flags(inputs.get('img_feats'), inputs.get('pts_feats'))      # the eager step's draw, on the host
        flags = (int(m.c_flag), int(m.l_flag))
        self._stage(inputs)
        entry = self.graphs.get(flags)
        if entry is None:
            entry = self._capture(flags)
        g, loss = entry
        g.replay()
        self.replayed_launches += self.launches[flags]
        if self.exchange == 'after':
            self.buckets.exchange_all()
            self.opt.step()
        return loss
