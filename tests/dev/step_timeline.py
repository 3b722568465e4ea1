"""dev: kernel timeline of MeasureVAE train steps (torch.profiler / CUPTI) -> gpurun_out/step_timeline.json
(per kernel: name, stream, start us, duration us) for gap analysis on the host without a GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
from inpaintnet_b200.measure_vae import MeasureVAE
from inpaintnet_b200.trainer import VAETrainer
from inpaintnet_b200.data import SyntheticFolkDataset

V, B = 64, 4096
ds = SyntheticFolkDataset(num_notes=V)
torch.manual_seed(0)
model = MeasureVAE(ds).cuda().set_precision("bf16")
tr = VAETrainer(ds, model, lr=1e-4)
model.train()
tok = torch.randint(0, V, (B, 24)).cuda()


def step(tf):
    model.decoder.teacher_forcing_prob = 2.0 if tf else -1.0
    tr.zero_grad()
    loss, acc = tr.loss_and_acc_for_batch(tok, 0, train=True)
    loss.backward()
    tr.step()


for i in range(6):
    step(i % 2 == 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(4):
        step(i % 2 == 0)
    torch.cuda.synchronize()
out = os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "gpurun_out", "step_timeline_trace.json")
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
rows = [dict(name=e["name"][:80], stream=e.get("args", {}).get("stream"), ts=e["ts"], dur=e["dur"])
        for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
rows.sort(key=lambda r: r["ts"])
json.dump(rows, open(out.replace("_trace.json", ".json"), "w"))
os.remove(out)
print(len(rows), "device events")
