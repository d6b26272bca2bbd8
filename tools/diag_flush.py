import time, torch
torch.cuda.init()
d = "cuda:0"
bufs = {"u8": torch.empty(256 << 20, dtype=torch.uint8, device=d),
        "f32": torch.empty(64 << 20, dtype=torch.float32, device=d),
        "f64": torch.empty(32 << 20, dtype=torch.float64, device=d)}
for name, b in bufs.items():
    for op in ("zero_", "fill1", "add"):
        ts = []
        for _ in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            if op == "zero_": b.zero_()
            elif op == "fill1": b.fill_(1)
            else: b.add_(1)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        print(name, op, " ".join(f"{t:.3f}" for t in ts))
