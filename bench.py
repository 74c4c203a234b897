#!/usr/bin/env python
"""bench.py -- the FPL+ hot path on B200: DSBN 3-D U-Net train step (voxels/s) + filtered-pseudo-label
pass (volumes/s), with the roofline of the dominant kernel and the reference's CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  Workload of the headline metric (BASELINE.json configs[2]; configs[1] is
reported in the same line under "pl_filter"): one ``training_all`` optimiser step of the final FPL+
segmentor = zero_grad, forward of a source batch (domain 0) and a pseudo-labelled target batch
(domain 1, pixel/image-weighted), 0.5*Dice+0.5*CE, backward, Adam (weight_decay 1e-5); batch 4 per
domain per GPU of 1x32x128x128 patches, UNet2D5_dsbn ft 16-32-64-128-256, 2 classes.  Synthetic
data of that shape, synthetic weights of that architecture.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

NET_PARAMS = {"net_type": "UNet2D5_dsbn", "num_domains": 2, "class_num": 2, "in_chns": 1,
              "feature_chns": [16, 32, 64, 128, 256], "conv_dims": [3, 3, 3, 3, 3],
              "dropout": [0.0, 0.0, 0.3, 0.4, 0.5], "bilinear": False, "deep_supervise": False, "aes": False}
PATCH = (32, 128, 128)
BATCH = 4                         # per domain per GPU (BASELINE.json configs[2])
VOLUME = (48, 256, 256)           # configs[1]
TRAIN_CFG = {"train_fpl_uda": True, "dual": True, "dis": False, "val_t1": False, "val_t2": False, "gpus": [0],
             "loss_type": ["DiceLoss", "CrossEntropyLoss"], "loss_weight": [0.5, 0.5], "optimizer": "Adam",
             "learning_rate": 1e-4, "momentum": 0.9, "weight_decay": 1e-5, "lr_scheduler": "MultiStepLR",
             "lr_gamma": 0.5, "lr_milestones": [10000, 20000, 30000], "iter_start": 0, "iter_max": 40000,
             "iter_valid": 500, "ckpt_save_dir": "/tmp/fplplus_bench", "deterministic": True, "random_seed": 1}
TEST_CFG = {"fpl": True, "gpus": [0], "domian_label": 1, "ae": False, "ckpt_mode": 2, "evaluation_mode": True,
            "test_time_dropout": True, "tta_mode": 1, "sliding_window_enable": True,
            "sliding_window_size": [32, 128, 128], "sliding_window_stride": [32, 128, 128]}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# --------------------------------------------------------------------------------------------
# synthetic data (synthetic_data.py is the seeded generator shared with the parity tests; it is data
# generation, not the measured path, and not part of oracle/)
# --------------------------------------------------------------------------------------------
def make_batch(seed, n, shape, weighted, pinned):
    import synthetic_data as synth
    x = torch.from_numpy(synth.synth_image(n, 1, shape, seed=seed))
    lab = synth.synth_label(n, 2, shape, seed=seed)
    d = {"image": x, "label_prob": torch.from_numpy(synth.one_hot(lab, 2))}
    if weighted:
        pw, iw = synth.synth_pixel_weight(lab, seed=seed)
        d["pixel_weight"] = torch.from_numpy(pw)
        d["image_weight"] = torch.from_numpy(iw)
    if pinned:
        d = {k: (v.pin_memory() if v.dtype == torch.float32 else v) for k, v in d.items()}
    return d


def batch_bytes(b):
    return sum(v.numel() * v.element_size() for k, v in b.items() if k != "image_weight")


# --------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# --------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thr = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# per-kernel timing through the C-ABI call hook: CUDA events on the launching stream
# --------------------------------------------------------------------------------------------
def _conv_work(a, off):
    n, d, h, w, cin, cout, kd = a[off:off + 7]
    return 2.0 * n * d * h * w * cin * cout * kd * 9, float(n * d * h * w * (cin + cout) * 2)


def _dfold_work(a):
    n, d, h, w, cin, cout = a[9:15]
    return 2.0 * n * d * h * w * cin * cout * 27, float(n * d * h * w * (cin + cout) * 2)


WORK = {   # entry point -> (flops, algorithmic bytes) from its argument tuple
    "fpl_conv3d_tc": lambda a: _conv_work(a, 9),
    "fpl_conv3d_tc_dfold": _dfold_work,
    "fpl_conv3d_direct": lambda a: _conv_work(a, 9),
    "fpl_conv3d_wgrad": lambda a: _conv_work(a, 7),
    "fpl_conv3d_wgrad_tc": lambda a: _conv_work(a, 7),
    "fpl_conv3d_wgrad_tc_tapmajor": lambda a: _conv_work(a, 7),
}


def _dsbn_work(bytes_per_elem, lo, hi):
    def f(a):
        n, d, h, w, c = a[lo:hi]
        return 0.0, float(n) * d * h * w * c * bytes_per_elem
    return f


# HBM-bound kernels: algorithmic bytes per element of DESIGN.md section 3 (bf16 activations)
WORK.update({
    "fpl_dsbn_bn_act_fwd": _dsbn_work(4, -6, -1),            # read y, write a
    "fpl_dsbn_act_bwd_reduce": _dsbn_work(4, -6, -1),        # read y, g
    "fpl_dsbn_act_bwd_apply_fin": _dsbn_work(6, -10, -5),    # read y, g, write dy
})
HBM_KERNELS = ("fpl_dsbn_bn_act_fwd", "fpl_dsbn_act_bwd_reduce", "fpl_dsbn_act_bwd_apply_fin")


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, averaged over the launches of one train step,
    from the committed `ncu --set full` capture (profiles/roofline_traffic.json, written by tools/ncu_traffic.py)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(kernel, {}).get("dram_bytes_per_launch")


class KernelTimer(object):
    class _Tok(object):
        __slots__ = ("rec", "e1")

        def __init__(self, rec, e1):
            self.rec, self.e1 = rec, e1

        def stop(self):
            self.e1.record()

    def __init__(self):
        self.records = []
        self.enabled = False

    def __call__(self, name, args):
        if not self.enabled:
            return None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        work = WORK[name](args) if name in WORK else (0.0, 0.0)
        e0.record()
        self.records.append((name, work, e0, e1))
        return KernelTimer._Tok(None, e1)

    def summary(self):
        agg = {}
        for name, (fl, by), e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += ms
            a[2] += fl
            a[3] += by
        return agg


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def build_agent(stage, world):
    from fplplus_b200.agent import GradAllReducer, SegmentationAgent
    import synthetic_data as synth
    cfg = {"dataset": {"tensor_type": "float", "train_batch_size": BATCH}, "network": dict(NET_PARAMS),
           "training": dict(TRAIN_CFG), "testing": dict(TEST_CFG)}
    agent = SegmentationAgent(cfg, stage)
    agent.create_network()
    sd = synth.synth_state_dict()
    agent.net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    agent._pick_device("training" if stage == "train" else "testing")
    agent.net.to(agent.device)
    if stage == "train":
        agent.create_optimizer(agent.get_parameters_to_update())
        agent.create_loss_calculator()
        agent.net.train()
        if world > 1:
            agent.reducer = GradAllReducer()
            agent.net.grad_ready_hook = agent.reducer.hook
            agent.net.grad_wait_hook = agent.reducer.finish
    return agent


def timed(fn, steps, warmup, barrier):
    """W warm-up calls, then exactly K calls between barrier+synchronize, CUDA events; ms total."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    timed.host_enqueue_ms = (time.time() - t0) * 1e3 / max(steps, 1)   # CPU time to enqueue one step
    torch.cuda.synchronize()
    barrier()
    t1 = time.time()
    return e0.elapsed_time(e1), t0, t1


def run_ours(args):
    import torch.distributed as dist
    from fplplus_b200 import fpl, lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    agent = build_agent("train", world)
    dev = agent.device
    host = [make_batch(11 + rank * 2, BATCH, PATCH, False, True), make_batch(12 + rank * 2, BATCH, PATCH, True, True)]
    resident = [{k: (v.to(dev) if torch.is_tensor(v) and v.dtype == torch.float32 else v) for k, v in b.items()}
                for b in host]
    vox_per_step = 2 * BATCH * PATCH[0] * PATCH[1] * PATCH[2]

    clocks = ClockSampler(local)
    clocks.start()

    # ---- kernel-side throughput: inputs resident in HBM; the step is replayed from the agent's CUDA graph ----
    def step_resident():
        agent.train_step(resident)

    for _ in range(max(args.warmup, 5)):          # includes the 3 eager steps + capture of the graph
        step_resident()
    torch.cuda.synchronize()
    ms, t0, t1 = timed(step_resident, args.steps, 0, barrier)
    host_ms = timed.host_enqueue_ms
    ms = max_over_ranks(ms)
    clk = clocks.stop(t0, t1)
    value = world * vox_per_step * args.steps / (ms / 1e3)

    # ---- end to end through the agent with HOST buffers: H2D of the batch + D2H of the loss per step ----
    # the loss of every step is read back to the host (4 bytes into pinned memory); the read of step i is
    # completed while step i+1 is already enqueued, so the GPU never idles on the host round trip
    loss_pin = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "last": None}

    def step_e2e():
        i = state["i"]
        loss, _ = agent.train_step(host)
        loss_pin[i & 1].copy_(loss, non_blocking=True)
        loss_ev[i & 1].record()
        if i > 0:
            loss_ev[(i - 1) & 1].synchronize()
            state["last"] = float(loss_pin[(i - 1) & 1])
        state["i"] = i + 1

    ms_e2e, _, _ = timed(step_e2e, args.steps, 2, barrier)
    ms_e2e = max_over_ranks(ms_e2e)
    e2e_value = world * vox_per_step * args.steps / (ms_e2e / 1e3)
    h2d = sum(batch_bytes(b) for b in host)

    # ---- roofline of the dominant kernel: a second timed region of K EAGER steps (a graph replay has no per-kernel
    #      host hook) with CUDA events on the launching stream around every C-ABI call; same kernels, same shapes ----
    timer = KernelTimer()
    lib.set_call_timer(timer)
    agent.use_cuda_graph = False
    # one stream for this pass: with the domain passes / wgrads on side streams an event bracket would also time the
    # kernels it overlaps with
    saved_streams = (agent.dual_stream, agent.net.wgrad_side_stream)
    agent.dual_stream, agent.net.wgrad_side_stream = False, False
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize()
    lib.launch_count(reset=True)
    timer.enabled = True
    ms_eager, _, _ = timed(step_resident, args.steps, 0, barrier)
    timer.enabled = False
    launches = lib.launch_count()
    agent.use_cuda_graph = True
    agent.dual_stream, agent.net.wgrad_side_stream = saved_streams
    pk = peaks()
    agg = timer.summary()
    step_ms = ms / args.steps
    eager_step_ms = ms_eager / args.steps
    kern = {k: {"launches_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps,
                "share_of_step": v[1] / args.steps / eager_step_ms} for k, v in agg.items()}
    roofline = None
    conv = {k: v for k, v in agg.items() if k in WORK and v[1] > 0 and v[2] > 0}
    if conv:
        top = max(conv, key=lambda k: conv[k][1])
        n, tot_ms, fl, by = conv[top]
        ach = fl / (tot_ms / 1e3) / 1e12
        roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"], "traffic": measured_traffic(top),
                    "traffic_note": "bytes per launch (dram read + write, ncu --set full, mean over the launches of one step)",
                    "peak_source": pk["source"] + " (sustained bf16, kernel timed inside a long step)",
                    "launches_per_step": n / args.steps, "avg_launch_ms": tot_ms / n,
                    "algorithmic_gflop_per_launch": fl / n / 1e9, "algorithmic_mb_per_launch": by / n / 1e6,
                    "hbm_gbs_at_algorithmic_bytes": by / (tot_ms / 1e3) / 1e9}
    # HBM-bound kernels: GB/s at the algorithmic bytes, over all layers of the step and for the largest layer alone
    hbm = {}
    for name in HBM_KERNELS:
        recs = [(by, e0.elapsed_time(e1)) for nm, (fl, by), e0, e1 in timer.records if nm == name and by > 0]
        if recs:
            big = max(b for b, _ in recs)
            tb = [t for b, t in recs if b == big]
            hbm[name] = {"gbs_all_layers": sum(b for b, _ in recs) / (sum(t for _, t in recs) / 1e3) / 1e9,
                         "gbs_largest_layer": big / (sum(tb) / len(tb) / 1e3) / 1e9,
                         "frac_of_hbm_peak_largest_layer": big / (sum(tb) / len(tb) / 1e3) / 1e9 / pk["hbm_gbs"],
                         "largest_layer_mb": big / 1e6}
    lib.set_call_timer(None)

    # ---- configs[1]: filtered-pseudo-label pass, volumes sharded over ranks, no communication ----
    pl = None
    if not args.skip_filter:
        pl = bench_filter(agent, rank, world, barrier, max_over_ranks, args)

    out = None
    if rank == 0:
        out = {"metric": "train_voxels_per_s", "value": value, "unit": "voxels/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": "FPL+ final segmentor training_all step: UNet2D5_dsbn ft 16-32-64-128-256, "
                                      "2 classes, batch 4/domain/GPU x 1x32x128x128, source batch + pixel/image-"
                                      "weighted target batch, 0.5 Dice + 0.5 CE, Adam (BASELINE.json configs[2])",
                          "voxels_per_step_per_gpu": vox_per_step, "parallelism": "dp%d" % world,
                          "l2": "activation working set per step >> 126 MB L2 (inputs larger than L2)",
                          "conv_gflop_per_step_per_gpu": 2 * BATCH * 179.9},
               "e2e": {"value": e2e_value, "unit": "voxels/s", "ms_per_step": ms_e2e / args.steps,
                       "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                       "api": "fplplus_b200.agent.SegmentationAgent.train_step(host batch dicts); every step's loss is read "
                              "back (pinned D2H), completed with a one-step lag"},
               "gpu_launches": launches, "gpu_launches_note": "kernels of this library launched by %d eager steps (the "
               "timed region replays the same kernels from a CUDA graph; torch's fused Adam adds 2 more per step)" % args.steps,
               "host_enqueue_ms_per_step": host_ms, "eager_ms_per_step": eager_step_ms,
               "clocks": clk,
               "roofline": roofline,
               "hbm_kernels": hbm,
               "kernels": kern,
               "conv_tensor_util": {"achieved_tflops_over_step": world * 2 * BATCH * 179.9e9 / (step_ms / 1e3) / 1e12 / world,
                                    "peak_tflops": pk["bf16_tflops_sustained"]},
               "pl_filter": pl}
        if not args.skip_cpu:
            out["cpu_baseline"] = cpu_baseline(bounded_steps=12)
    if out is not None:
        print(json.dumps(out), flush=True)
    if world > 1:
        # captured graphs hold NCCL work: drop them before the communicator, and do not let a wedged
        # communicator teardown turn a finished measurement into a hang
        dist.barrier()
        torch.cuda.synchronize()
        agent._graphs.clear()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def bench_filter(agent, rank, world, barrier, max_over_ranks, args):
    """Full FPL filter pass per volume: target pass (domain 1) + fake-source pass (domain 0) ->
    labels + agreement pixel weights; K=6 MC-dropout passes -> image uncertainty.  8 Inferer runs x
    (8 windows x 4 flips) = 256 network forwards of 1x32x128x128 per 48x256x256 volume."""
    from fplplus_b200 import fpl
    from fplplus_b200.inferer import Inferer
    import synthetic_data as synth
    import torch.nn as nn
    net = agent.net
    net.eval()
    cfg = dict(TEST_CFG)
    cfg["class_num"] = 2
    inferer = Inferer(cfg)
    nvol = max(1, args.volumes)
    vols = [torch.from_numpy(synth.synth_image(1, 1, VOLUME, seed=50 + rank * nvol + i)).pin_memory() for i in range(nvol)]
    fake = [torch.from_numpy(synth.synth_image(1, 1, VOLUME, seed=150 + rank * nvol + i)).pin_memory() for i in range(nvol)]
    results = []

    def one_volume(i):
        with torch.no_grad():
            tgt = vols[i].to(agent.device, non_blocking=True)
            src = fake[i].to(agent.device, non_blocking=True)
            one = torch.ones(1, dtype=torch.long)
            for m in net.modules():
                if type(m) == nn.Dropout:
                    m.eval()
            z_t = inferer.run(net, tgt, 1 * one)
            z_s = inferer.run(net, src, 0 * one)
            la, lb, w, cnt = fpl.agreement_weight(z_t, z_s)
            for m in net.modules():
                if type(m) == nn.Dropout:
                    m.train()
            passes = inferer.run(net, tgt, 1 * one, mc_passes=6)
            stats, _ = fpl.mc_uncertainty(passes)
            lab_host = la.cpu()                         # what leaves the GPU: u8 labels, fp32 weights, 2 scalars
            w_host = w.cpu()
            u = fpl.finish_uncertainty(stats)
            results.append((lab_host.shape, w_host.shape, u))

    one_volume(0)                                       # warm-up volume
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(nvol):
        one_volume(i)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1))
    net.train()
    vps = world * nvol / (ms / 1e3)
    return {"metric": "pl_filter_volumes_per_s", "value": vps, "unit": "volumes/s", "ms_per_volume": ms / nvol,
            "volumes_timed_per_gpu": nvol, "forwards_per_volume": 256,
            "conv_tflops": 15.35 * vps / world,
            "workload": "VS-style 1x48x256x256 volume: dual-domain sliding-window inference (window 32x128x128, "
                        "4-flip TTA) + argmax labels + agreement pixel weights + 6 MC-dropout passes -> image "
                        "uncertainty (BASELINE.json configs[1]); host volumes in, u8 labels + fp32 weights + scalar out",
            "scaling": "weak (volumes sharded round-robin, no collective)"}


# --------------------------------------------------------------------------------------------
# CPU: the oracle port of the reference path (the reference is pure Python/torch; its agent cannot
# be imported without SimpleITK/tensorboardX, see oracle/__init__.py) on the host cores
# --------------------------------------------------------------------------------------------
def _oracle_trainer(threads):
    import synthetic_data as synth
    from oracle.train_step import OracleTrainer
    torch.set_num_threads(threads)
    params = dict(NET_PARAMS)
    return OracleTrainer(synth.synth_state_dict(), params, lr=1e-4, weight_decay=1e-5, w_dice=0.5, w_ce=0.5)


def cpu_baseline(bounded_steps=2, batch=1):
    threads = os.cpu_count() or 1
    tr = _oracle_trainer(threads)
    b0, b1 = make_batch(11, batch, PATCH, False, False), make_batch(12, batch, PATCH, True, False)
    batches = [(b0["image"], b0["label_prob"], None), (b1["image"], b1["label_prob"], b1["pixel_weight"])]
    tr.step(batches)
    t0 = time.time()
    for _ in range(bounded_steps):
        tr.step(batches)
    dt = (time.time() - t0) / bounded_steps
    vox = 2 * batch * PATCH[0] * PATCH[1] * PATCH[2]
    return {"value": vox / dt, "unit": "voxels/s", "cores": threads, "kind": "port",
            "sample": "%d training_all steps (1 warm-up) at batch %d/domain of 1x32x128x128 (1/%d of the GPU arm's "
                      "step), fp32 torch CPU (oneDNN), oracle.train_step.OracleTrainer" % (bounded_steps, batch, BATCH // batch),
            "s_per_step": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    tr = _oracle_trainer(threads)
    batch = 1
    b0, b1 = make_batch(11, batch, PATCH, False, False), make_batch(12, batch, PATCH, True, False)
    batches = [(b0["image"], b0["label_prob"], None), (b1["image"], b1["label_prob"], b1["pixel_weight"])]
    for _ in range(max(1, min(args.warmup, 2))):
        tr.step(batches)
    t0 = time.time()
    for _ in range(args.steps):
        tr.step(batches)
    dt = time.time() - t0
    vox = 2 * batch * PATCH[0] * PATCH[1] * PATCH[2]
    v = vox * args.steps / dt
    sample = ("each step = one training_all step at batch 1/domain of 1x32x128x128 (1/4 of the GPU arm's per-GPU step), "
              "fp32 torch CPU on all host threads; warm-up capped at 2 steps")
    out = {"impl": "reference", "metric": "train_voxels_per_s", "value": v, "unit": "voxels/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "FPL+ final segmentor training_all step (BASELINE.json configs[2]) on the host CPU, "
                                  "bounded sample", "parallelism": "cpu"},
           "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    wd = int(os.environ.get("FPL_BENCH_WATCHDOG", "0"))
    if wd > 0:                       # debugging aid: dump every thread's stack and exit if the run wedges
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--volumes", type=int, default=2, help="volumes timed per GPU in the pl_filter leg")
    ap.add_argument("--skip-filter", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    run_ours(args)


if __name__ == "__main__":
    main()
