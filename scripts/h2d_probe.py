"""Host-to-device copy bandwidth per GPU with 1, 2, 4, 8 GPUs copying at once (pinned 512 MiB buffers, 1 s each):
is the end-to-end ceiling of an 8-GPU box 8 x one PCIe link, or a shared host fabric?
    python scripts/h2d_probe.py > gpurun_out/h2d_probe.json"""
import json, multiprocessing as mp, os, sys, time


def worker(rank, k, barrier, q):
    import torch
    torch.cuda.set_device(rank)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    host = torch.empty(512 << 20, dtype=torch.uint8).pin_memory()
    dev = torch.empty_like(host, device="cuda")
    dev.copy_(host, non_blocking=True); torch.cuda.synchronize()
    barrier.wait()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t_end = time.perf_counter() + 1.0
    e0.record()
    while time.perf_counter() < t_end:
        for _ in range(4):
            dev.copy_(host, non_blocking=True); n += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    q.put((rank, n * host.numel() / (e0.elapsed_time(e1) / 1e3) / 1e9))


def main():
    import torch
    ng = torch.cuda.device_count()
    out = {"gpus_visible": ng, "runs": []}
    ctx = mp.get_context("spawn")
    for k in (1, 2, 4, 8):
        if k > ng:
            break
        barrier, q = ctx.Barrier(k), ctx.Queue()
        ps = [ctx.Process(target=worker, args=(r, k, barrier, q)) for r in range(k)]
        [p.start() for p in ps]
        res = sorted(q.get() for _ in range(k))
        [p.join() for p in ps]
        per = [round(v, 2) for _, v in res]
        out["runs"].append({"gpus_copying": k, "gb_per_s_per_gpu": per, "gb_per_s_total": round(sum(per), 1)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
