#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps 30 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1     # CPU arm: the UNMODIFIED reference

Headline workload = BASELINE.json configs[1]: MeasureVAE training step, bf16 tensor-core mode, batch 4096 synthetic
24-tick measures per GPU (weak scaling), reference default hyper-parameters (V=64, E=10, H=512, L=2, Z=256,
dropout 0.5), forward + fused CE/KL loss + backward + gradient all-reduce (N>1) + fused Adam inside every timed
step.  The reference flips one teacher-forcing coin per batch (MeasureVAE/decoder.py:431-434) and the two outcomes
run different kernels, so the timed steps ALTERNATE teacher-forced / argmax (exactly 50/50 for even K) and the two
modes are also timed on their own (`modes`).  The same JSON line carries the other configs the metric names:
`inpaint` (configs[3]: batched inpainting inference, queries/s), `latent_train` (configs[2]: LatentRNN training with
the frozen MeasureVAE, data parallel) and `arnn_train` (configs[4]: AnticipationRNN training step, 4096 sequences of
384 ticks per GPU).  One JSON line is printed by rank 0.
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
# SURVEY.md section 8(d) / BASELINE.md section 4: algorithmic FLOPs per unit at V=64
FLOPS_PER_MEASURE_TRAIN = 1.476e9
FLOPS_PER_QUERY_646 = 4.85e9          # 6/4/6 inpainting query, unused target-encode skipped (6.10e9 with it)
FLOPS_PER_SEQ_ARNN_TRAIN = 4.48e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="measures per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--inpaint-queries", type=int, default=8192, help="inpainting queries per GPU (0 = skip)")
    ap.add_argument("--latent-seqs", type=int, default=256, help="LatentRNN training: 16-measure sequences per GPU per step (0 = skip)")
    ap.add_argument("--arnn-seqs", type=int, default=4096, help="AnticipationRNN training: 384-tick sequences per GPU per step (0 = skip)")
    ap.add_argument("--sections", default="mvae,inpaint,latent,arnn")
    # CPU arm
    ap.add_argument("--workload", default="mvae_train", choices=["mvae_train", "inpaint", "latent_train", "arnn_train"])
    ap.add_argument("--cpu-size", type=int, default=0, help="units per step of the CPU arm (0 = sized to the time budget)")
    ap.add_argument("--cpu-budget-s", type=float, default=120.0)
    ap.add_argument("--cpu-extra", default="", help="json kwargs of the CPU workload")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference on the host cores (oracle/ref_bench.py); the oracle port only when the
# reference is neither mounted nor staged
# ---------------------------------------------------------------------------------------------
CPU_DEFAULT_SIZE = {"mvae_train": 1024, "inpaint": 256, "latent_train": 32, "arnn_train": 32}
CPU_METRIC = {"mvae_train": (METRIC, UNIT), "inpaint": ("inpaint_queries_per_sec", "queries/s"),
              "latent_train": ("sequences_per_sec_latent_rnn_train_step", "sequences/s"),
              "arnn_train": ("sequences_per_sec_arnn_train_step", "sequences/s")}


class CpuPort:
    """Fallback: the oracle's restatement of the MeasureVAE train step (kind "port")."""

    def __init__(self, batch, seed=0):
        from oracle import inpaintnet_oracle as O
        from tests.golden import recipe
        self.O = O
        torch.manual_seed(seed)
        random.seed(seed)
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


def time_cpu(workload, size, steps, warmup, budget_s, extra):
    """-> dict(value, unit, cores, kind, sample, s_per_step, units_per_step)."""
    from oracle import ref_bench
    if not ref_bench.available():
        if workload != "mvae_train":
            return None
        cores = ref_bench._threads()
        size = size or 256
        port = CpuPort(size)
        for _ in range(warmup):
            port.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            port.step()
        per = (time.perf_counter() - t0) / steps
        return dict(value=size / per, unit=UNIT, cores=cores, kind="port", s_per_step=per, units_per_step=size,
                    sample=f"{steps} train steps of {size} measures after {warmup} warm-up (oracle port: the reference is "
                           f"neither mounted nor staged in oracle/_ref), fp32, {per:.2f} s/step")
    if not size:
        # bounded sample: the largest power-of-two fraction of the default size whose (steps + warmup) fit the budget,
        # estimated from one probe step at 1/4 of the default size (CPU throughput is flat from there on: BASELINE.md)
        full = CPU_DEFAULT_SIZE[workload]
        probe = ref_bench.run(workload, max(1, full // 4), 1, 0, **extra)
        size = full
        while size > max(1, full // 4) and probe["s_per_step"] * (size / (full // 4)) * (steps + warmup) > budget_s:
            size //= 2
    return ref_bench.run(workload, size, steps, warmup, **extra)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    extra = json.loads(args.cpu_extra) if args.cpu_extra else {}
    if not extra.get("cuda"):
        # the CPU arm: the reference's helpers move tensors to the GPU whenever one is visible (utils/helpers.py:5-26)
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
        torch.cuda.is_available = lambda: False
    r = time_cpu(args.workload, args.cpu_size, steps, warmup, args.cpu_budget_s, extra)
    metric, unit = CPU_METRIC[args.workload]
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "reference not staged (run oracle/make_ref.sh)"}))
        return
    line = {
        "impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": {"mvae_train": "MeasureVAE train step (BASELINE.json configs[1]), CPU bounded sample",
                                "inpaint": "InpaintNet batched inference (BASELINE.json configs[3]), CPU bounded sample",
                                "latent_train": "LatentRNN train step (BASELINE.json configs[2]), CPU bounded sample",
                                "arnn_train": "AnticipationRNN train step (BASELINE.json configs[4]), CPU bounded sample"}[args.workload],
                   "units_per_step": r["units_per_step"], "V": V, "E": E, "H": H, "Z": Z, "layers": 2},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if extra.get("cuda"):
        line["config"]["workload"] = line["config"]["workload"].replace("CPU bounded sample", "reference's own torch/cuDNN GPU path")
        line["dtype"] = "f32 (cuDNN)"
    print(json.dumps(line))


def cpu_baseline_subprocess(workload, steps, warmup, size=0, extra=None, budget_s=60.0):
    """Runs the CPU arm in its own interpreter (its own thread pool, no module-name clash with the drop-in packages)."""
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "WORLD_SIZE",
                                                            "LOCAL_RANK", "CUDA_VISIBLE_DEVICES")}
    if not (extra or {}).get("cuda"):
        env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", str(steps),
           "--warmup", str(warmup), "--cpu-size", str(size), "--cpu-budget-s", str(budget_s)]
    if extra:
        cmd += ["--cpu-extra", json.dumps(extra)]
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                d = json.loads(ln)
                return d.get("cpu_baseline") or {"unavailable": d.get("unavailable")}
        return {"unavailable": "no JSON line from the CPU arm: " + out.stderr[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}


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
        # median over the samples taken UNDER LOAD (the sampler also sees the idle gaps between sections)
        busy = sorted(x for x in sm if smax is None or x >= 0.6 * smax) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# helpers shared by the sections
# ---------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch.distributed as dist
        self.args = args
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        peaks, src = {}, "measured (MEASURED_PEAKS.json: sustained bf16 matmul / HBM copy)"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            src = "fallback (B200_PROFILING.md: 1.4 PFLOP/s sustained bf16, 6.65 TB/s)"
        self.tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        self.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_src = src
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            self.traffic = {}

    def sync_all(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def timed(self, fn, steps, finish=None):
        """CUDA events on the current stream around `steps` calls, barrier + synchronize on both sides, max over ranks."""
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        e1.record()
        self.sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def kernel_table(self, kernels):
        return {n: {"launches": k["launches"], "ms": round(k["ms"], 4),
                    "tflops": round(k["flops"] / (k["ms"] * 1e-3) / 1e12, 2) if k["ms"] > 0 and k["flops"] > 0 else None,
                    "gbs": round(k["bytes"] / (k["ms"] * 1e-3) / 1e9, 1) if k["ms"] > 0 and k["bytes"] > 0 else None}
                for n, k in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])}

    def roofline_of(self, name, k, total_ms):
        """Roofline entry of one kernel class from the CUDA-event timings of ipn_prof_* (launching stream)."""
        per_launch_ms = k["ms"] / max(1, k["launches"])
        tensor = k["flops"] > 0
        if tensor:
            ach = k["flops"] / (k["ms"] * 1e-3) / 1e12
            r = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": self.tf_peak, "unit": "TFLOP/s", "frac": ach / self.tf_peak}
        else:
            ach = k["bytes"] / (k["ms"] * 1e-3) / 1e9
            r = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": self.hbm_peak, "unit": "GB/s", "frac": ach / self.hbm_peak}
        r.update({"traffic": self.traffic.get(name), "launch_us": per_launch_ms * 1e3, "launches": k["launches"],
                  "share_of_kernel_time": k["ms"] / total_ms, "peak_source": self.peak_src})
        if tensor and k["bytes"] > 0:
            r["hbm_gbs"] = k["bytes"] / (k["ms"] * 1e-3) / 1e9
        return r

    def rooflines(self, kernels, critical=()):
        total = sum(k["ms"] for k in kernels.values()) or 1.0
        if not kernels:
            return None, []
        name, k = max(kernels.items(), key=lambda kv: kv[1]["ms"])
        crit = [self.roofline_of(n, kernels[n], total) for n in critical if n in kernels]
        return self.roofline_of(name, k, total), crit

    def profile(self, fn, n):
        """Per-kernel-class CUDA-event breakdown of n calls of fn (rank 0 records; every rank runs: collectives)."""
        from inpaintnet_b200 import ops
        if self.rank == 0:
            ops.prof_enable(True)
        for i in range(n):
            fn(i)
        torch.cuda.synchronize()
        kernels = {}
        if self.rank == 0:
            kernels = ops.prof_report()
            ops.prof_enable(False)
        if self.world > 1:
            self.dist.barrier()
        return kernels


def standalone_wgrad(cx, B, reps=20):
    """The dominant weight-gradient GEMM of the step -- dW_hh[1536, 512] += dP[24*B, 1536]^T . h[24*B, 512] of one
    encoder-layer direction -- timed ALONE on an idle GPU (CUDA events, `reps` back-to-back launches, operands 403 MB:
    larger than the L2).  Inside the step these GEMMs run on a side stream next to the persistent chain kernels, which
    hold 128 of the 148 SMs, so their in-step launch durations (the `roofline` entry) mostly measure waiting for SMs;
    this is the kernel's own rate."""
    from inpaintnet_b200 import ops
    rows, H3, Hh = 24 * B, 3 * H, H
    dP = torch.randn(rows, H3, device="cuda").bfloat16()
    hp = torch.randn(rows, Hh, device="cuda").bfloat16()
    g = torch.zeros(H3, Hh, device="cuda")

    def one(_i):
        ops.gemm(ops.CORE_UMMA, ops.BF16, H3, Hh, [(dP.data_ptr(), H3, 1, hp.data_ptr(), Hh, 1, rows)], g.data_ptr(), ops.F32, Hh,
                 accumulate=ops.ATOMIC_ADD)

    for i in range(3):
        one(i)
    ms = cx.timed(one, reps) / reps
    tf = 2.0 * rows * H3 * Hh / (ms * 1e-3) / 1e12
    return {"kernel": "gemm_umma_tn_wgrad", "shape": f"dW[{H3},{Hh}] += dP[{rows},{H3}]^T h[{rows},{Hh}]", "launch_us": ms * 1e3,
            "achieved": tf, "unit": "TFLOP/s", "peak": cx.tf_peak, "frac": tf / cx.tf_peak,
            "how": f"{reps} launches alone on an idle GPU, CUDA events; algorithmic bytes {(rows * (H3 + Hh) * 2) / 1e6:.0f} MB per launch"}


# ---------------------------------------------------------------------------------------------
# section: MeasureVAE training step (configs[1]) -- the headline metric
# ---------------------------------------------------------------------------------------------
def run_mvae(cx, K, W):
    from inpaintnet_b200 import ops
    from inpaintnet_b200.measure_vae import MeasureVAE
    from inpaintnet_b200.trainer import VAETrainer, LaggedReadback
    from inpaintnet_b200.data import SyntheticFolkDataset
    args, world, rank = cx.args, cx.world, cx.rank
    torch.manual_seed(0)
    random.seed(0)
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
    dec = model.decoder
    coin = {"mix": lambda i: 2.0 if i % 2 == 0 else -1.0, "teacher_forced": lambda i: 2.0, "argmax": lambda i: -1.0}
    mode = ["mix"]

    def step_resident(i):
        dec.teacher_forcing_prob = coin[mode[0]](i)      # forced coin: exactly alternating in "mix"
        trainer.zero_grad()
        loss, acc = trainer.loss_and_acc_for_batch(dev_tokens[i % 4], 0, train=True)
        loss.backward()
        trainer.step()
        return loss

    readback = LaggedReadback()
    e2e_losses = []

    def step_e2e(i):
        # the trainer's own per-batch call (what Trainer.loss_and_acc_on_epoch runs): async H2D of the pinned int32
        # host batch, the step, and an async D2H of (loss, accuracy, guard flags) that the host reads one step later
        dec.teacher_forcing_prob = coin["mix"](i)
        trainer.run_batch((host_batches[i % 4], None), 0, train=True, readback=readback)
        e2e_losses.extend(readback.pop(keep=1))

    def finish_e2e():
        e2e_losses.extend(readback.pop(keep=0))   # waits for the last step's results: inside the timed region

    for i in range(W):
        step_resident(i)
    l0 = ops.launch_count()
    ms = cx.timed(step_resident, K)
    launches = ops.launch_count() - l0
    modes = {"mix": {"ms_per_step": ms / K, "value": world * B * K / (ms / 1e3)}}
    for m in ("teacher_forced", "argmax"):
        mode[0] = m
        step_resident(0)
        ms_m = cx.timed(step_resident, K)
        modes[m] = {"ms_per_step": ms_m / K, "value": world * B * K / (ms_m / 1e3)}
    mode[0] = "mix"
    for i in range(W):          # the end-to-end path has its own one-time costs (pinned result slots, first async copies)
        step_e2e(i)
    finish_e2e()
    e2e_losses.clear()
    ms_e2e = cx.timed(step_e2e, K, finish_e2e)
    assert len(e2e_losses) == K and all(l == l for l, _ in e2e_losses), "e2e arm: every step's loss must reach the host"
    trainer.check_device_flags()
    kernels = cx.profile(step_resident, max(2, args.profile_steps) // 2 * 2)   # equal numbers of TF and argmax steps
    dec.teacher_forcing_prob = 0.5
    standalone = standalone_wgrad(cx, B)
    out = dict(B=B, K=K, ms=ms, ms_e2e=ms_e2e, launches=launches, modes=modes, kernels=kernels, standalone=standalone)
    del trainer, model
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
# section: batched inpainting inference (configs[3])
# ---------------------------------------------------------------------------------------------
def run_inpaint(cx):
    from inpaintnet_b200.measure_vae import MeasureVAE
    from inpaintnet_b200.latent_rnn import LatentRNN
    from inpaintnet_b200.data import SyntheticFolkDataset
    args, world, rank = cx.args, cx.world, cx.rank
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
    dev = host.cuda().long()
    out_host = torch.empty(Q, 1, 24 * 4, dtype=torch.int64).pin_memory()

    def split(score, n_t):
        n_p = (16 - n_t) // 2          # 6/4/6: script_gen_diff_models.py:144-146 ; 7/2/7: test_reconstruction.py:52
        return score[:, :n_p], score[:, n_p:n_p + n_t], score[:, n_p + n_t:]

    def q_resident(n_t=4, with_target=False):
        def fn(_i):
            past, target, future = split(dev, n_t)
            with torch.no_grad():
                if with_target:    # the reference also encodes the target measures and discards them (latent_rnn.py:133)
                    model.vae_model.encoder(target.reshape(-1, 24))
                return model(past, future, target, n_t, train=False)
        return fn

    def q_e2e(_i):
        score = host.cuda(non_blocking=True).long()
        past, target, future = split(score, 4)
        with torch.no_grad():
            w, s, z = model(past, future, target, 4, train=False)
        out_host.copy_(s, non_blocking=True)   # the decoded tokens come back to the host
        torch.cuda.current_stream().synchronize()

    # the same call captured once in a CUDA graph (inpaintnet_b200.inference.GraphedInpainter: one launch per batch
    # instead of ~250; fresh context-latent noise per replay as in the eager call)
    from inpaintnet_b200.inference import GraphedInpainter
    dev32 = host.cuda()
    gi = GraphedInpainter(model, Q, 6, 4, 6)

    def q_graph_resident(_i):
        return gi(dev32)

    def q_graph_e2e(_i):
        w, s, z = gi(host)
        out_host.copy_(s, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    reps = 5
    res = {}
    for name, fn in (("resident", q_resident()), ("e2e", q_e2e), ("n_target_2", q_resident(2)),
                     ("incl_target_encode", q_resident(4, True)), ("graph_resident", q_graph_resident),
                     ("graph_e2e", q_graph_e2e)):
        for i in range(3):
            fn(i)
        res[name] = cx.timed(fn, reps) / reps
    kernels = cx.profile(q_resident(), 1)
    qps = lambda ms: world * Q / (ms / 1e3)
    out = {"metric": "inpaint_queries_per_sec", "unit": "queries/s", "value": qps(res["graph_resident"]), "queries_per_gpu": Q,
           "split": "6/4/6", "ms_per_batch": res["graph_resident"], "reps": reps, "warmup": 3,
           "achieved_model_tflops": qps(res["graph_resident"]) * FLOPS_PER_QUERY_646 / 1e12,
           "e2e": {"value": qps(res["graph_e2e"]), "unit": "queries/s", "ms_per_batch": res["graph_e2e"], "h2d_bytes_per_step": Q * 384 * 4,
                   "d2h_bytes_per_step": Q * 96 * 8,
                   "how": "GraphedInpainter(model, Q, 6, 4, 6)(pinned host int32 tokens): async H2D into the graph's input, fresh "
                          "Philox noise, ONE graph launch (= LatentRNN.forward(past, future, target, 4, train=False): encode 12 "
                          "context measures, context + generation GRUs, argmax decode of 4 gap measures), D2H of the int64 tokens"},
           "eager": {"value": qps(res["resident"]), "ms_per_batch": res["resident"], "e2e_value": qps(res["e2e"]),
                     "e2e_ms_per_batch": res["e2e"], "how": "the same LatentRNN.forward call issued launch by launch (~250 launches)"},
           "variants": {"n_target_2_split_7_2_7": {"value": qps(res["n_target_2"]), "ms_per_batch": res["n_target_2"]},
                        "incl_unused_target_encode": {"value": qps(res["incl_target_encode"]), "ms_per_batch": res["incl_target_encode"]}},
           "note": "value = inputs resident in HBM, CUDA-graph replay of the eager call; target-encode (unused by the "
                   "non-autoregressive model) skipped; `eager` and `variants` are eager-mode timings"}
    del gi
    if rank == 0:
        roof, crit = cx.rooflines(kernels, critical=("gru_layer_fwd_persist", "tick_decode_persist", "gru_step_fwd_umma"))
        out["roofline"], out["roofline_critical_path"], out["kernels"] = roof, crit, cx.kernel_table(kernels)
    del model, vae
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
# section: LatentRNN training with the frozen MeasureVAE (configs[2])
# ---------------------------------------------------------------------------------------------
def run_latent_train(cx, K, W):
    from inpaintnet_b200.measure_vae import MeasureVAE
    from inpaintnet_b200.latent_rnn import LatentRNN
    from inpaintnet_b200.trainer import LatentRNNTrainer, LaggedReadback
    from inpaintnet_b200.data import SyntheticFolkDataset
    args, world, rank = cx.args, cx.world, cx.rank
    S = args.latent_seqs
    ds = SyntheticFolkDataset(num_notes=V)
    out = {"metric": "sequences_per_sec_latent_rnn_train_step", "unit": "sequences/s", "sequences_per_gpu_per_step": S,
           "measures_per_sequence": 16, "split": "stochastic, latent_rnn_trainer.py:77-132, rank-shared seed", "modes": {}}
    g = torch.Generator().manual_seed(4321 + rank)
    host_batches = [torch.randint(0, V, (S, 1, 384), generator=g, dtype=torch.int32).pin_memory() for _ in range(4)]
    for name, auto_reg in (("auto_reg_false", False), ("auto_reg_true_teacher_forcing", True)):
        torch.manual_seed(0)
        random.seed(0)
        vae = MeasureVAE(ds)
        model = LatentRNN(ds, vae, 2, 512, 0.5, torch.nn.GRU, auto_reg=auto_reg, teacher_forcing=True)
        model.cuda()
        model.set_precision(args.precision)
        trainer = LatentRNNTrainer(ds, model, lr=1e-4)
        model.train()
        if auto_reg:
            model.teacher_forcing_prob = 2.0   # the script default (train_inpaintnet.py:53-56) with the coin on its TF side
        readback = LaggedReadback()
        losses = []

        def step(i):   # the trainer's own per-batch call on a pinned host batch (split drawn on the host, rank-shared)
            trainer.run_batch((host_batches[i % 4], None), 0, train=True, readback=readback)
            losses.extend(readback.pop(keep=1))

        def finish():
            losses.extend(readback.pop(keep=0))

        torch.manual_seed(7)     # rank-shared: every rank draws the same past/gap/future split (same shapes, same kernels)
        for i in range(W):
            step(i)
        finish()
        losses.clear()
        torch.manual_seed(11)
        ms = cx.timed(step, K, finish)
        assert len(losses) == K and all(l == l for l, _ in losses)
        trainer.check_device_flags()
        m = {"value": world * S * K / (ms / 1e3), "measures_per_s": world * S * 16 * K / (ms / 1e3), "ms_per_step": ms / K,
             "steps": K, "warmup": W, "h2d_bytes_per_step": S * 384 * 4, "d2h_bytes_per_step": 16}
        if not auto_reg:
            torch.manual_seed(11)
            kernels = cx.profile(step, 2)
            finish()
            if rank == 0:
                m["roofline"], m["roofline_critical_path"] = cx.rooflines(kernels, critical=("gru_layer_fwd_persist", "gru_layer_bwd_persist",
                                                                                           "gru_step_fwd_umma", "gru_step_bwd_umma"))
                m["kernels"] = cx.kernel_table(kernels)
        out["modes"][name] = m
        del trainer, model, vae
        torch.cuda.empty_cache()
    out["value"] = out["modes"]["auto_reg_false"]["value"]
    out["ms_per_step"] = out["modes"]["auto_reg_false"]["ms_per_step"]
    out["note"] = ("value = auto_reg=False (the model the evaluation scripts load, test_reconstruction.py:141); end to end through "
                   "Trainer.run_batch on pinned host batches; frozen VAE from a fixed-seed init; dropout on (model.train() "
                   "reaches the frozen VAE as in the reference)")
    return out


# ---------------------------------------------------------------------------------------------
# section: AnticipationRNN training step (configs[4])
# ---------------------------------------------------------------------------------------------
def run_arnn_train(cx, K, W):
    from inpaintnet_b200.arnn import ConstraintModelGaussianReg, AnticipationRNNGaussianRegTrainer
    from inpaintnet_b200.trainer import LaggedReadback
    from inpaintnet_b200.data import SyntheticFolkDataset
    args, world, rank = cx.args, cx.world, cx.rank
    S = args.arnn_seqs
    ds = SyntheticFolkDataset(num_notes=V)
    torch.manual_seed(0)
    random.seed(0)
    model = ConstraintModelGaussianReg(dataset=ds, note_embedding_dim=10, metadata_embedding_dim=2, num_lstm_constraints_units=256,
                                       num_lstm_generation_units=256, linear_hidden_size=256, num_layers=2, dropout_input_prob=0.2,
                                       dropout_prob=0.2, unary_constraint=True, teacher_forcing=True)   # train_arnn_reg.py:13-26,86-98
    model.cuda()
    model.set_precision(args.precision)
    trainer = AnticipationRNNGaussianRegTrainer(ds, model, lr=1e-4)
    model.train()
    g = torch.Generator().manual_seed(999 + rank)
    score = torch.randint(0, V, (S, 1, 384), generator=g, dtype=torch.int32).pin_memory()
    t = torch.arange(384)
    meta = torch.stack([((t // 6) % 4 == 0).long(), t % 6, torch.zeros_like(t)], 1).to(torch.int32)
    meta = meta.view(1, 1, 384, 3).expand(S, 1, 384, 3).contiguous().pin_memory()
    readback = LaggedReadback()
    losses = []

    def step(i):
        trainer.run_batch((score, meta), 0, train=True, readback=readback)
        losses.extend(readback.pop(keep=1))

    def finish():
        losses.extend(readback.pop(keep=0))

    out = {"metric": "sequences_per_sec_arnn_train_step", "unit": "sequences/s", "sequences_per_gpu_per_step": S, "ticks": 384,
           "lstm_hidden": 256, "layers": "2 constraint + 2 generation", "modes": {}}
    Ka, Wa = max(2, min(K, 6)), max(1, min(W, 2))
    for name, prob in (("teacher_forced", 2.0), ("no_teacher_forcing", -1.0)):
        model.teacher_forcing_prob = prob
        torch.manual_seed(5)     # rank-shared gap location (anticipation_rnn_trainer.py:93-128)
        for i in range(Wa):
            step(i)
        finish()
        losses.clear()
        torch.manual_seed(9)
        ms = cx.timed(step, Ka, finish)
        assert len(losses) == Ka and all(l == l for l, _ in losses)
        m = {"value": world * S * Ka / (ms / 1e3), "ms_per_step": ms / Ka, "steps": Ka, "warmup": Wa,
             "achieved_model_tflops": world * S * Ka / (ms / 1e3) * FLOPS_PER_SEQ_ARNN_TRAIN / 1e12,
             "h2d_bytes_per_step": S * 384 * 4 * 4, "d2h_bytes_per_step": 16}
        torch.manual_seed(9)
        kernels = cx.profile(step, 1)
        finish()
        losses.clear()
        if rank == 0:
            m["roofline"], m["roofline_critical_path"] = cx.rooflines(
                kernels, critical=("lstm_layer_fwd_persist", "lstm_layer_bwd_persist", "lstm_step_fwd_umma", "lstm_step_bwd_umma"))
            m["kernels"] = cx.kernel_table(kernels)
            m["launches_per_step"] = sum(k["launches"] for k in kernels.values())
        out["modes"][name] = m
    trainer.check_device_flags()
    out["value"] = out["modes"]["teacher_forced"]["value"]
    out["ms_per_step"] = out["modes"]["teacher_forced"]["ms_per_step"]
    out["note"] = ("value = teacher-forced branch; end to end through Trainer.run_batch on pinned host (score, metadata) batches; "
                   "the reference flips a coin per batch (arnn_model.py:425-428), both branches are reported")
    del trainer, model
    torch.cuda.empty_cache()
    return out


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

    cx = Ctx(args)
    dist, world, rank, local = cx.dist, cx.world, cx.rank, cx.local
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a B200: inpaintnet_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, (world, args.gpus)
    sections = set(args.sections.split(","))
    W, K = max(3, args.warmup), max(1, args.steps)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    mv = run_mvae(cx, K, W)
    clocks = sampler.stop() if rank == 0 else None
    inpaint = run_inpaint(cx) if "inpaint" in sections and args.inpaint_queries > 0 else None
    latent = run_latent_train(cx, max(2, min(K, 10)), 3) if "latent" in sections and args.latent_seqs > 0 else None
    arnn = run_arnn_train(cx, K, W) if "arnn" in sections and args.arnn_seqs > 0 else None

    if rank == 0:
        B, ms, ms_e2e = mv["B"], mv["ms"], mv["ms_e2e"]
        value = world * B * K / (ms / 1e3)
        e2e_value = world * B * K / (ms_e2e / 1e3)
        kernels = mv["kernels"]
        roofline, crit = cx.rooflines(kernels, critical=("gru_layer_fwd_persist", "gru_layer_bwd_persist", "tick_decode_persist"))
        if roofline is not None:
            roofline["standalone"] = mv["standalone"]
            roofline["note"] = ("achieved = in-step launch durations (side stream, next to chain kernels that hold 128 SMs); "
                                "standalone = the same kernel alone on the GPU")
        # the serial chain: per-step latency of one encoder-layer step against the floors of DESIGN.md section 4.3
        # (12.9 GFLOP at the sustained tensor peak; 84 MB of algorithmic bytes at the measured HBM rate)
        for c in crit:
            if c["kernel"] == "tick_decode_persist":
                c["floor_note"] = ("argmax decode, 24 serial ticks of 4096 measures in one launch: 0.47 TFLOP -> %.0f us at the tensor "
                                   "peak; a tick is a chain of 4 MMA phases, 3 epilogues and 2 cluster exchanges (DESIGN.md section 4.5)"
                                   % (0.47e12 / (cx.tf_peak * 1e12) * 1e6))
                continue
            c["floor_note"] = ("encoder layer step at B=4096, both directions: 12.9 GFLOP -> %.1f us at the tensor peak; 84 MB "
                               "algorithmic bytes -> %.1f us at the HBM peak" % (12.9e9 / (cx.tf_peak * 1e12) * 1e6,
                                                                                84e6 / (cx.hbm_peak * 1e9) * 1e6))
        cpu = ref_gpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_subprocess("mvae_train", 3, 1, size=1024)
            # context only (SURVEY.md section 2.2 "the bar"): the UNMODIFIED reference on this same B200 through torch's
            # cuDNN GRU path, fp32, at the full 4096-measure batch, with the reference's own ~50 host syncs per step
            ref_gpu = cpu_baseline_subprocess("mvae_train", 5, 2, size=B, extra={"cuda": True})
            if inpaint is not None:
                inpaint["cpu_baseline"] = cpu_baseline_subprocess("inpaint", 2, 1, size=256)
            if latent is not None:
                latent["cpu_baseline"] = cpu_baseline_subprocess("latent_train", 3, 1, size=32)
            if arnn is not None:
                arnn["cpu_baseline"] = cpu_baseline_subprocess("arnn_train", 3, 1, size=32)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "MeasureVAE train step, BASELINE.json configs[1]", "measures_per_gpu_per_step": B,
                       "V": V, "E": E, "H": H, "Z": Z, "layers": 2, "dropout": 0.5,
                       "teacher_forcing": "coin p=0.5 per batch in the reference; timed steps alternate teacher-forced / argmax",
                       "parallelism": f"dp{world}", "l2": "working set per step (several GB) exceeds the 126 MB L2",
                       "achieved_model_tflops": value * FLOPS_PER_MEASURE_TRAIN / 1e12},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 24 * 4, "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / K,
                    "how": "Trainer.run_batch per step: pinned int32 tokens -> async H2D, step, async D2H of "
                           "(loss, accuracy, 2 guard flags) read by the host one step later; last read inside the timed region"},
            "gpu_launches": int(mv["launches"]),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_critical_path": crit,
            "cpu_baseline": cpu,
            "reference_on_this_gpu": ref_gpu,
            "modes": mv["modes"],
            "kernels": cx.kernel_table(kernels),
            "inpaint": inpaint,
            "latent_train": latent,
            "arnn_train": arnn,
        }
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
