"""dev: MeasureVAE train step with the batch pipelined as n micro-batches on n streams (VAETrainer.microbatches)."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from inpaintnet_b200 import ops
from inpaintnet_b200.measure_vae import MeasureVAE
from inpaintnet_b200.trainer import VAETrainer
from inpaintnet_b200.data import SyntheticFolkDataset

torch.manual_seed(0); random.seed(0)
B, V = int(os.environ.get("B", 4096)), 64
ds = SyntheticFolkDataset(num_notes=V)
model = MeasureVAE(ds); model.cuda(); model.set_precision("bf16"); model.train()
tr = VAETrainer(ds, model, lr=1e-4)
dev = [torch.randint(0, V, (B, 24)).cuda() for _ in range(4)]

def step(i):
    tr.zero_grad()
    loss, acc = tr.loss_and_acc_for_batch(dev[i % 4], 0, train=True)
    loss.backward()
    tr.step()
    return loss

for n, cap in ((1, 0), (2, 0), (2, 116), (2, 84), (4, 0), (4, 100), (1, 0)):
    tr.microbatches = n
    ops.GEMM_MAIN_MAX_CTAS = cap
    for mode, prob in (("tf", 2.0), ("argmax", -1.0), ("coin", 0.5)):
        model.decoder.teacher_forcing_prob = prob
        random.seed(1)
        for i in range(3): step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for i in range(10): l = step(i)
        t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
        print(f"microbatches={n} gemm_cap={cap:3d} {mode:6s} {e0.elapsed_time(e1) / 10:7.2f} ms/step  host issue {(t1 - t0) * 100:6.2f} ms/step  loss {l.item():.4f}", flush=True)
