"""PCIe probe (development aid): pinned H2D / D2H bandwidth of the box, to judge the e2e number."""
import json
import subprocess
import time

import torch

dev = torch.device("cuda:0")
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(n, dtype=torch.uint8, device=dev)
out = {}
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    out[name + "_GBps"] = 4 * n / (time.perf_counter() - t0) / 1e9
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s1):
        d2.copy_(h2, non_blocking=True)
    with torch.cuda.stream(s2):
        h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out["duplex_each_GBps"] = 4 * n / dt / 1e9
try:
    q = subprocess.check_output(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current",
                                 "--format=csv,noheader", "-i", "0"], text=True).strip()
    out["pcie"] = q
except Exception as ex:
    out["pcie"] = str(ex)
# host memset bandwidth (single thread) for the host-side ts fill estimate
t0 = time.perf_counter(); h2.fill_(1); out["host_fill_GBps_1thread"] = n / (time.perf_counter() - t0) / 1e9
print(json.dumps(out))
