import sys, os, json, subprocess, tempfile
sys.path.insert(0, os.getcwd())
import bench
iq_u, ps_u, U = bench.make_input(2048, 2048)
for i in range(4):
    r = bench.brick_leg(iq_u)
    print(i, r["value"], r["frames_per_s"], flush=True)
print(os.cpu_count(), len(os.sched_getaffinity(0)), open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else "-")
