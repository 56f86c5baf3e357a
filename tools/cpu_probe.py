import multiprocessing as mp, os, sys, time
sys.path.insert(0, ".")
import bench
def burn(_):
    t = time.perf_counter(); x = 0
    while time.perf_counter() - t < 2.0: x += 1
    return x
if __name__ == "__main__":
    print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
        if os.path.exists(p): print(p, open(p).read().strip())
    os.system("lscpu | grep -E 'Model name|^CPU\\(s\\)|Thread|Core|Socket' ")
    ctx = mp.get_context("spawn")
    for n in (1, 2, 4, 8, 16):
        with ctx.Pool(n) as pool:
            t = time.perf_counter(); r = pool.map(burn, range(n)); dt = time.perf_counter() - t
        print("burn workers", n, "iters/worker %.2e" % (sum(r) / n), "wall %.2f" % dt)
    for n in (1, 4, 16):
        with ctx.Pool(n) as pool:
            t = time.perf_counter(); r = pool.map(bench.cpu_pass, [(1000 + w, 400, 200) for w in range(n)]); dt = time.perf_counter() - t
        print("oracle workers", n, "opt sec per worker", ["%.2f" % s for s, _ in r][:4], "trials", r[0][1], "wall %.1f" % dt)
