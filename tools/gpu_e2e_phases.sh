#!/bin/bash
# host-side phases of the end-to-end step (stage / profile / gather) + the raw pinned H2D / D2H rate of the box
python - <<'PY'
import torch, time
x = torch.empty(400 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(x, device='cuda')
for name, a, b in (("h2d", d, x), ("d2h", x, d)):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); print(name, "GB/s", round(5 * 0.4194 / (time.perf_counter() - t), 1))
PY
MEPP_BENCH_DEBUG=1 timeout 600 python bench.py --workload ${1:-cfg3} --steps 10 --warmup 3 --no-cpu-baseline --no-accuracy --no-clocks --no-fullrun --no-tensor-leg > gpurun_out/e2e_phases.json 2> gpurun_out/e2e_phases.err
grep "e2e host ms" gpurun_out/e2e_phases.err | cut -c 1-1500
python - <<'PY'
import json
d = json.loads(open("gpurun_out/e2e_phases.json").read().strip().splitlines()[-1])
print("value ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), d["details"]["host_ms_per_step"])
PY
