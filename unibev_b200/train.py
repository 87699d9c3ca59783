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
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
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
