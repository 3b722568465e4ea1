#!/usr/bin/env python
"""bench.py -- MeasureVAE train-step throughput (measures/s) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps 30 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1     # CPU arm (oracle port)

Workload = BASELINE.json configs[1]: MeasureVAE training step, bf16 tensor-core mode, batch 4096
synthetic 24-tick measures per GPU (weak scaling), reference default hyper-parameters (V=64, E=10,
H=512, L=2, Z=256, dropout 0.5, teacher-forcing coin p=0.5 per step as MeasureVAE/decoder.py:431-434),
forward + fused CE/KL loss + backward + gradient all-reduce (N>1) + fused Adam inside every timed step.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

V, E, H, Z = 64, 10, 512, 256
METRIC = "measures_per_sec_mvae_train_step"
UNIT = "measures/s"
FLOPS_PER_MEASURE_TRAIN = 1.476e9  # algorithmic fwd+bwd FLOPs per measure at V=64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="measures per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=256, help="measures per step of the CPU arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--inpaint-queries", type=int, default=8192, help="inpainting queries per GPU (0 = skip)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's train step (the reference itself is Python under
# /root/reference and does not exist on the GPU box)
# ---------------------------------------------------------------------------------------------
class CpuPort:
    def __init__(self, batch, seed=0):
        from oracle import inpaintnet_oracle as O
        from tests.golden import recipe
        self.O = O
        torch.manual_seed(seed)
        random.seed(seed)
        torch.set_num_threads(os.cpu_count() or 1)
        sd = recipe.make_state_dict(recipe.mvae_spec(V, E, H, Z), seed)
        self.params = {k: v.clone().requires_grad_() for k, v in sd.items()}
        self.m = {k: torch.zeros_like(v) for k, v in sd.items()}
        self.v = {k: torch.zeros_like(v) for k, v in sd.items()}
        self.batch = batch
        self.tokens = torch.randint(0, V, (batch, 24))
        self.step_no = 0

    def step(self):
        O, B = self.O, self.batch
        self.step_no += 1
        for p in self.params.values():
            p.grad = None
        keep = dict(enc=[(torch.rand(B, 24, 2 * H) > 0.5).float()], beat=[(torch.rand(B, 4, H) > 0.5).float()],
                    tick=(torch.rand(B, 24, H) > 0.5).float())
        tf = random.random() < 0.5
        w, s, mu, ls, z = O.mvae_forward(self.params, self.tokens, torch.randn(B, Z), tf, train_dropout=keep)
        loss = O.mvae_loss(w, self.tokens, mu, ls)
        loss.backward()
        with torch.no_grad():
            for k, p in self.params.items():
                O.adam_step(p, p.grad, self.m[k], self.v[k], self.step_no)
        return loss.item()


def time_cpu_port(batch, steps, warmup):
    port = CpuPort(batch)
    for _ in range(warmup):
        port.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        port.step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    value, per_step = time_cpu_port(args.cpu_batch, steps, warmup)
    cores = torch.get_num_threads()
    sample = f"{steps} train steps of {args.cpu_batch} measures (fwd+bwd+Adam, dropout on, TF coin), oracle port, fp32"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MeasureVAE train step (BASELINE.json configs[1]), CPU bounded sample",
                   "measures_per_step": args.cpu_batch, "V": V, "E": E, "H": H, "Z": Z, "layers": 2},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_inpaint(args, world, rank, timed):
    from inpaintnet_b200.measure_vae import MeasureVAE
    from inpaintnet_b200.latent_rnn import LatentRNN
    from inpaintnet_b200.data import SyntheticFolkDataset
    Q = args.inpaint_queries
    ds = SyntheticFolkDataset(num_notes=V)
    torch.manual_seed(0)
    vae = MeasureVAE(ds)
    model = LatentRNN(ds, vae, 2, 512, 0.5, torch.nn.GRU, auto_reg=False)
    model.cuda()
    model.set_precision(args.precision)
    model.eval()
    g = torch.Generator().manual_seed(99 + rank)
    host = torch.randint(0, V, (Q, 16, 24), generator=g, dtype=torch.int32).pin_memory()
    n_p, n_t, n_f = 6, 4, 6   # script_gen_diff_models.py:144-146

    def query(_i):
        score = host.cuda(non_blocking=True).long()
        past, target, future = score[:, :n_p], score[:, n_p:n_p + n_t], score[:, n_p + n_t:]
        with torch.no_grad():
            w, s, z = model(past, future, target, n_t, train=False)
        return s.cpu()   # the decoded tokens come back to the host

    for i in range(2):
        query(i)
    reps = 5
    ms = timed(query, reps)
    kern = None
    if rank == 0:   # per-kernel-class breakdown of one query batch
        from inpaintnet_b200 import ops
        ops.prof_enable(True)
        query(0)
        kern = {n: {"launches": k["launches"], "ms": round(k["ms"], 4)} for n, k in
                sorted(ops.prof_report().items(), key=lambda kv: -kv[1]["ms"])}
        ops.prof_enable(False)
    return {"metric": "inpaint_queries_per_sec", "value": world * Q * reps / (ms / 1e3), "unit": "queries/s",
            "queries_per_gpu": Q, "split": "6/4/6", "ms_per_batch": ms / reps, "end_to_end": True, "kernels": kern,
            "note": "host int32 tokens -> H2D -> encode 12 context measures + LatentRNN + argmax decode of 4 gap "
                    "measures -> D2H tokens; target-encode (unused by the non-autoregressive model) skipped"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    # stdout carries exactly ONE JSON line: python-level prints (model constructors) and C-level ones (NCCL's
    # version banner) are sent to stderr until the final line is written
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved_fd, "w")

    import torch.distributed as dist
    from inpaintnet_b200 import ops
    from inpaintnet_b200.measure_vae import MeasureVAE
    from inpaintnet_b200.trainer import VAETrainer
    from inpaintnet_b200.data import SyntheticFolkDataset

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a B200: inpaintnet_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, (world, args.gpus)

    torch.manual_seed(0)
    random.seed(0)  # rank-shared: every rank flips the same teacher-forcing coin (same kernels, balanced step)
    B = args.batch
    ds = SyntheticFolkDataset(num_notes=V)
    model = MeasureVAE(ds)  # reference defaults
    model.cuda()
    model.set_precision(args.precision)
    trainer = VAETrainer(ds, model, lr=1e-4)
    model.train()
    g = torch.Generator().manual_seed(1234 + rank)
    seqs = B // 16
    host_batches = [torch.randint(0, V, (seqs, 1, 384), generator=g, dtype=torch.int32).pin_memory() for _ in range(4)]
    dev_tokens = [hb.view(seqs * 16, 24).long().cuda() for hb in host_batches]

    def step_resident(i):
        trainer.zero_grad()
        loss, acc = trainer.loss_and_acc_for_batch(dev_tokens[i % 4], 0, train=True)
        loss.backward()
        trainer.step()
        return loss

    from inpaintnet_b200.trainer import LaggedReadback
    readback = LaggedReadback()
    e2e_losses = []

    def step_e2e(i):
        # the trainer's own per-batch call (what Trainer.loss_and_acc_on_epoch runs): async H2D of the pinned int32
        # host batch, the step, and an async D2H of (loss, accuracy, guard flags) that the host reads one step later
        trainer.run_batch((host_batches[i % 4], None), 0, train=True, readback=readback)
        e2e_losses.extend(readback.pop(keep=1))

    def finish_e2e():
        e2e_losses.extend(readback.pop(keep=0))   # waits for the last step's results: inside the timed region

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    W, K = max(3, args.warmup), max(1, args.steps)
    for i in range(W):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count()
    random.seed(77)   # both timed arms flip the same coins; this seed gives exactly 50 % teacher-forced steps at K = 10, 30, 100 (a TF step is ~2.5 ms shorter)
    ms = timed(step_resident, K)
    launches = ops.launch_count() - l0
    for i in range(W):          # the end-to-end path has its own one-time costs (pinned result slots, first async copies)
        step_e2e(i)
    finish_e2e()
    e2e_losses.clear()
    random.seed(77)
    ms_e2e = timed(step_e2e, K, finish_e2e)
    assert len(e2e_losses) == K and all(l == l for l, _ in e2e_losses), "e2e arm: every step's loss must reach the host"
    clocks = sampler.stop() if rank == 0 else None
    trainer.check_device_flags()
    value = world * B * K / (ms / 1e3)
    e2e_value = world * B * K / (ms_e2e / 1e3)

    # ---- per-kernel-class breakdown with CUDA events on the launching stream (roofline numbers)
    kernels = {}
    if rank == 0:
        ops.prof_enable(True)
    random.seed(1)
    for i in range(max(2, args.profile_steps)):   # every rank steps (the gradient all-reduce is collective)
        model.decoder.teacher_forcing_prob = 2.0 if i % 2 == 0 else -1.0   # one TF step, one argmax step
        step_resident(i)
    model.decoder.teacher_forcing_prob = 0.5
    if rank == 0:
        kernels = ops.prof_report()
        ops.prof_enable(False)
    if world > 1:
        dist.barrier()

    # ---- second half of BASELINE.json's metric: batched inpainting inference (configs[3]): encode the 12
    # context measures, LatentRNN generates 4 gap latents, argmax decode; queries sharded over ranks.
    inpaint = None
    if args.inpaint_queries > 0:
        inpaint = run_inpaint(args, world, rank, timed)

    if rank == 0:
        peaks = {}
        src = "measured (MEASURED_PEAKS.json, sustained bf16 / HBM copy)"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            src = "fallback (B200_PROFILING.md: 1.4 PFLOP/s sustained bf16, 6.65 TB/s)"
        tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        total_ms = sum(k["ms"] for k in kernels.values()) or 1.0
        top = max(kernels.items(), key=lambda kv: kv[1]["ms"]) if kernels else (None, None)
        roofline = None
        if top[0] is not None:
            name, k = top
            per_launch_ms = k["ms"] / k["launches"]
            if k["flops"] > 0:
                ach = k["flops"] / k["launches"] / (per_launch_ms * 1e-3) / 1e12
                roofline = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                            "frac": ach / tf_peak, "traffic": None, "launch_us": per_launch_ms * 1e3,
                            "share_of_step": k["ms"] / total_ms, "peak_source": src}
            else:
                ach = k["bytes"] / k["launches"] / (per_launch_ms * 1e-3) / 1e9
                roofline = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                            "frac": ach / hbm_peak, "traffic": None, "launch_us": per_launch_ms * 1e3,
                            "share_of_step": k["ms"] / total_ms, "peak_source": src}
            tr = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the ncu --set full capture
            if os.path.exists(tr):
                roofline["traffic"] = json.load(open(tr)).get(name)
        breakdown = {n: {"launches": k["launches"], "ms": round(k["ms"], 4),
                         "tflops": (k["flops"] / (k["ms"] * 1e-3) / 1e12) if k["ms"] > 0 and k["flops"] > 0 else None,
                         "gbs": (k["bytes"] / (k["ms"] * 1e-3) / 1e9) if k["ms"] > 0 and k["bytes"] > 0 else None}
                     for n, k in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cb = args.cpu_batch
            v, per = time_cpu_port(cb, 3, 1)
            cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"3 train steps of {cb} measures after 1 warm-up (oracle port of the reference step, fp32, "
                             f"{per:.2f} s/step)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "MeasureVAE train step, BASELINE.json configs[1]", "measures_per_gpu_per_step": B,
                       "V": V, "E": E, "H": H, "Z": Z, "layers": 2, "dropout": 0.5, "teacher_forcing": "coin p=0.5 per step",
                       "parallelism": f"dp{world}", "l2": "working set per step (several GB) exceeds the 126 MB L2",
                       "achieved_model_tflops": value * FLOPS_PER_MEASURE_TRAIN / 1e12},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 24 * 4, "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / K,
                    "how": "Trainer.run_batch per step: pinned int32 tokens -> async H2D, step, async D2H of "
                           "(loss, accuracy, 2 guard flags) read by the host one step later; last read inside the timed region"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "kernels": breakdown,
            "inpaint": inpaint,
        }
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
