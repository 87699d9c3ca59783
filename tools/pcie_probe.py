#!/usr/bin/env python
"""Host <-> device copy ceiling of the box with N ranks copying at once (what bounds bench.py's `e2e` at 4+ GPUs):
H2D only, D2H only and both directions, from ordinary pinned memory and from write-combined pinned memory.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py"""
import ctypes
import os

import torch
import torch.distributed as dist


def wc_pinned(nbytes):
    """cudaHostAlloc(..., cudaHostAllocWriteCombined) wrapped as a uint8 tensor (never freed: probe only)."""
    rt = ctypes.CDLL('libcudart.so.12', mode=ctypes.RTLD_GLOBAL) if False else None
    cudart = torch.cuda.cudart()
    ptr = ctypes.c_void_p()
    lib = None
    for name in ('libcudart.so.12', 'libcudart.so'):
        try:
            lib = ctypes.CDLL(name)
            break
        except OSError:
            continue
    if lib is None:
        import glob
        cand = glob.glob(os.path.join(os.path.dirname(torch.__file__), '..', 'nvidia', 'cuda_runtime', 'lib', 'libcudart.so*'))
        lib = ctypes.CDLL(cand[0])
    rc = lib.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(4))    # 4 = write-combined
    assert rc == 0, rc
    buf = (ctypes.c_ubyte * nbytes).from_address(ptr.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n = 168 << 20
    src = torch.empty(n, dtype=torch.uint8).pin_memory()
    dst = torch.empty(n, dtype=torch.uint8).pin_memory()
    src_wc = wc_pinned(n)
    src.fill_(1)
    src_wc.fill_(1)
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d_src, do_h2d, do_d2h, reps=8):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(reps):
            if do_h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h2d_src, non_blocking=True)
            if do_d2h:
                with torch.cuda.stream(s2):
                    dst.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nbytes = n * reps * (int(do_h2d) + int(do_d2h)) * world
        return nbytes / (float(ms.item()) / 1e3) / 1e9

    rows = [('H2D pinned', run(src, True, False)), ('H2D write-combined', run(src_wc, True, False)),
            ('D2H pinned', run(src, False, True)), ('both, pinned', run(src, True, True)),
            ('both, H2D write-combined', run(src_wc, True, True))]
    if rank == 0:
        for name, gbs in rows:
            print('%d ranks  %-26s %7.1f GB/s aggregate' % (world, name, gbs), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
