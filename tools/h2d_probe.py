"""Host -> device copy rate of the e2e arm's pyramid upload (103 MB): three level tensors vs one buffer,
alone and under a concurrent copy kernel / matmuls.  Run on the GPU box: python tools/h2d_probe.py
(measured: 55 GB/s for the three level tensors and under load; profiles/README.md)"""
import os, time, torch

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
shapes = [(5, 256, 128, 240), (5, 256, 64, 120), (5, 256, 32, 60)]
host = [torch.empty(s, dtype=torch.bfloat16).pin_memory() for s in shapes]
dst = [torch.empty(s, dtype=torch.bfloat16, device=dev) for s in shapes]
nbytes = sum(t.numel() * 2 for t in host)
one_h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
one_d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))


def rate(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    return ms, nbytes / ms / 1e6


def three():
    for d_, h_ in zip(dst, host):
        d_.copy_(h_, non_blocking=True)


def one():
    one_d.copy_(one_h, non_blocking=True)


print("three tensors: %.3f ms  %.1f GB/s" % rate(three))
print("one buffer   : %.3f ms  %.1f GB/s" % rate(one))
# under load: a bandwidth-heavy kernel on another stream
big_a = torch.empty(1 << 28, dtype=torch.bfloat16, device=dev)
big_b = torch.empty_like(big_a)
side = torch.cuda.Stream()


def loaded(fn):
    def g():
        with torch.cuda.stream(side):
            big_b.copy_(big_a)
        fn()
    return g


print("one buffer under an HBM copy kernel: %.3f ms  %.1f GB/s (includes the kernel if longer)" % rate(loaded(one)))
mm_a = torch.randn(4096, 4096, device=dev, dtype=torch.bfloat16)


def loaded_mm(fn):
    def g():
        with torch.cuda.stream(side):
            for _ in range(4):
                torch.matmul(mm_a, mm_a)
        fn()
    return g


print("one buffer under matmuls: %.3f ms  %.1f GB/s" % rate(loaded_mm(one)))
