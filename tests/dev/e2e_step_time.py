"""dev: where the end-to-end step loses time against the device-resident step (host upload / readback variants)."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from inpaintnet_b200.measure_vae import MeasureVAE
from inpaintnet_b200.trainer import VAETrainer, LaggedReadback
from inpaintnet_b200.data import SyntheticFolkDataset

import torch.distributed as dist
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
torch.manual_seed(0); random.seed(0)
B, V = 4096, 64
ds = SyntheticFolkDataset(num_notes=V)
model = MeasureVAE(ds); model.cuda(); model.set_precision("bf16"); model.train()
tr = VAETrainer(ds, model, lr=1e-4)
host = [torch.randint(0, V, (B // 16, 1, 384), dtype=torch.int32).pin_memory() for _ in range(4)]
dev = [h.view(B, 24).long().cuda() for h in host]

def core(tokens):
    tr.zero_grad()
    loss, acc = tr.loss_and_acc_for_batch(tokens, 0, train=True)
    loss.backward()
    tr.step()
    return loss, acc

def resident(i): core(dev[i % 4])
def upload_only(i): core(tr.process_batch_data((host[i % 4], None)))
def sync_read(i):
    loss, _ = core(tr.process_batch_data((host[i % 4], None)))
    return float(loss.detach().cpu())
rb = LaggedReadback()
def lagged(i):
    tr.run_batch((host[i % 4], None), 0, True, readback=rb); rb.pop(keep=1)
def lagged2(i):
    tr.run_batch((host[i % 4], None), 0, True, readback=rb); rb.pop(keep=2)
def lagged_push_only(i):
    tr.run_batch((host[i % 4], None), 0, True, readback=rb); rb.pending.clear()

def timed(fn, n=10, fin=None):
    for i in range(3): fn(i)
    if fin: fin()
    torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n): fn(i)
    t_issue = time.perf_counter() - t0
    if fin: fin()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_issue / n * 1e3

for name, fn, fin in (("resident", resident, None), ("upload_only", upload_only, None), ("sync_read", sync_read, None),
                      ("lagged", lagged, lambda: rb.pop(keep=0)), ("lagged2", lagged2, lambda: rb.pop(keep=0)),
                      ("lagged_push_only", lagged_push_only, None),
                      ("resident", resident, None)):
    ms, issue = timed(fn, fin=fin)
    print(f"[rank {rank}/{world} overlap={os.environ.get('IPN_DP_OVERLAP', '1')}] {name:18s} {ms:7.2f} ms/step   host issue {issue:6.2f} ms/step", flush=True)
if world > 1:
    dist.destroy_process_group()
