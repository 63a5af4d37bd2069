"""What HBM delivers on this box for the three access mixes the kernels have: copy (read + write, the mix
MEASURED_PEAKS.json's hbm_gbs is quoted on), write-only (the forward: 90 % of its DRAM bytes are the pooled-output
write) and read-only (backward phase 1: 97 % reads).  torch ops over 2 GiB buffers, best of 10, CUDA events.
usage: python scripts/hbm_rw_probe.py"""
import torch


def best(fn, reps=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    n = 1 << 29                                   # 2 GiB of fp32
    a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
    b = torch.empty_like(a)
    gb = n * 4 / 1e9
    t = best(lambda: b.copy_(a));        print(f"copy   (read+write) {2 * gb / t * 1e3:8.0f} GB/s  ({t:.3f} ms)")
    t = best(lambda: b.fill_(1.0));      print(f"fill   (write only) {gb / t * 1e3:8.0f} GB/s  ({t:.3f} ms)")
    t = best(lambda: b.zero_());         print(f"memset (write only) {gb / t * 1e3:8.0f} GB/s  ({t:.3f} ms)")
    t = best(lambda: a.sum());           print(f"sum    (read only)  {gb / t * 1e3:8.0f} GB/s  ({t:.3f} ms)")
    t = best(lambda: torch.add(a, 1.0, out=b)); print(f"add    (read+write) {2 * gb / t * 1e3:8.0f} GB/s  ({t:.3f} ms)")


if __name__ == "__main__":
    main()
