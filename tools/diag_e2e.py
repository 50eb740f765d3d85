"""Times the host<->device legs of the numpy call path (diagnostic; not part of the product)."""
import time

import numpy as np
import torch

n = 165636900
dev = torch.device("cuda", 0)
x = torch.zeros(n, dtype=torch.float64, device=dev)
torch.cuda.synchronize()


def t(label, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{label:45s} {best*1e3:8.1f} ms  {8*n/best/1e9:6.1f} GB/s", flush=True)
    return r


pin = t("torch.empty(pin_memory=True) alloc", lambda: torch.empty(n, dtype=torch.float64, pin_memory=True))
t("torch.empty(pin_memory=True) alloc (2nd buffer)", lambda: torch.empty(n, dtype=torch.float64, pin_memory=True))
t("D2H into pinned", lambda: pin.copy_(x))
t("H2D from pinned tensor", lambda: x.copy_(pin, non_blocking=True))
npv = pin.numpy()
t("from_numpy(pinned view).is_pinned()", lambda: torch.from_numpy(npv).is_pinned())
t("H2D from_numpy(pinned view).to(dev)", lambda: torch.from_numpy(npv).to(dev, non_blocking=True))
page = np.zeros(n)
t("H2D from pageable numpy", lambda: torch.from_numpy(page).to(dev))
t("D2H .cpu() pageable", lambda: x.cpu())
