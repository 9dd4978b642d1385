"""Per-kernel device-time breakdown of one full training step (torch.profiler / CUPTI; nsys is
not installed).  Prints the top kernels by total CUDA time and their share of the step."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unsloth_b200.ddp import FlatLoRABucket  # noqa: E402
from unsloth_b200.patch import build_qlora_model, lora_parameters  # noqa: E402


def main():
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda", 0)
    model = build_qlora_model("llama-3-8b", device=dev, num_hidden_layers=layers)
    bucket = FlatLoRABucket(lora_parameters(model))
    ids = torch.randint(0, 128256, (4, 2048), device=dev)

    def step():
        bucket.zero_grad()
        loss = model(input_ids=ids, labels=ids).loss
        loss.backward()
        bucket.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    tot = sum(e.device_time_total for e in ev)
    ev.sort(key=lambda e: -e.device_time_total)
    print(json.dumps({"layers": layers, "total_device_ms": round(tot / 1e3, 2)}))
    for e in ev[:40]:
        print(json.dumps({"kernel": e.key[:110], "calls": e.count, "ms": round(e.device_time_total / 1e3, 3),
                          "share": round(e.device_time_total / tot, 4)}))
    s, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); step(); e2.record(); torch.cuda.synchronize()
    print(json.dumps({"step_wall_ms_events": round(s.elapsed_time(e2), 2)}))


if __name__ == "__main__":
    main()
