#!/bin/bash
# Probe the GPU box: host cores/RAM, GPU, pinned-memory limits.
mkdir -p gpurun_out
{
echo "== nproc"; nproc
echo "== meminfo"; head -5 /proc/meminfo
echo "== ulimit -l"; ulimit -l
echo "== lscpu"; lscpu | head -25
echo "== nvidia-smi"; nvidia-smi
echo "== topo"; nvidia-smi topo -m
echo "== pcie"; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max --format=csv
echo "== df"; df -h /dev/shm /tmp . 2>/dev/null
echo "== hugepages"; grep -i huge /proc/meminfo
python - <<'PY'
import torch, time
print("torch", torch.__version__, torch.cuda.is_available(), torch.cuda.get_device_name(0))
p = torch.cuda.get_device_properties(0)
print(p)
for gb in (1, 8, 32):
    t=time.time(); x=torch.empty(gb*(1<<28), dtype=torch.float32, pin_memory=True); torch.cuda.synchronize(); dt=time.time()-t
    print(f"pin_memory alloc {gb} GiB: {dt:.2f}s")
    d=torch.empty(1<<28, dtype=torch.float32, device='cuda')
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    d.copy_(x[:1<<28], non_blocking=True); torch.cuda.synchronize()
    e0.record(); d.copy_(x[:1<<28], non_blocking=True); e1.record(); torch.cuda.synchronize()
    print(f"  H2D 1GiB: {1.073741824/(e0.elapsed_time(e1)/1e3):.1f} GB/s")
    e0.record(); x[:1<<28].copy_(d, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print(f"  D2H 1GiB: {1.073741824/(e0.elapsed_time(e1)/1e3):.1f} GB/s")
    del x
PY
} > gpurun_out/probe_box.txt 2>&1
tail -5 gpurun_out/probe_box.txt
