import torch, time
dev = torch.device("cuda:0")
for nbytes, label in ((2166671, "2 x 2.17 MB (one per camera)"), (4333342, "1 x 4.33 MB"), (8666684, "1 x 8.67 MB")):
    reps = 2 if "2 x" in label else 1
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(reps * 8)]
    devb = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(reps * 8)]
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(2):
        with torch.cuda.stream(s):
            e0.record()
            for k in range(200):
                for r in range(reps):
                    j = (k * reps + r) % len(host)
                    devb[j].copy_(host[j], non_blocking=True)
            e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{label}: {200 * reps * nbytes / ms / 1e6:.1f} GB/s, {ms / 200 * 1e3:.1f} us per window")
