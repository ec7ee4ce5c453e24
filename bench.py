#!/usr/bin/env python
"""bench.py - points/sec forward+backward of the GAPartNet sparse U-Net hot path on B200.

Workload (BASELINE.json configs[2], the one `metric` is quoted on): full sparse U-Net backbone
(in=6, channels [16,32,48,64,80,96,112], block_repeat 2) forward + backward on synthetic
20 000-point scenes, batch 16 per GPU, voxel 0.02.  One step = voxelize + all 13 rulebooks +
backbone forward + per-point gather + semantic head/CE loss + full backward (+ NCCL gradient
allreduce when N > 1).  Weak scaling: every rank processes its own batch of 16 scenes.

  python bench.py --gpus N --steps K --warmup W          (torchrun launches N ranks for N > 1)
  python bench.py --impl reference ...                   (CPU reference arm: the oracle port)

Prints ONE JSON line (rank 0).  `value` = points/s with inputs resident in HBM; `e2e` = the same
step driven from pinned HOST buffers (H2D of the points + D2H of the loss inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHANNELS = [16, 32, 48, 64, 80, 96, 112]
BLOCK_REPEAT = 2
IN_CH = 6
NUM_CLASSES = 10
PTS = 20000
BATCH = 16
VOXEL = 0.02
SHAPE = 128
METRIC = "points/sec fwd+bwd, 20k-pt scenes b16"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the L0 16->16 kernels, from the committed
# `ncu --set full` captures (profiles/r1_summary.md); null until captured
TRAFFIC_NCU = {"k_conv_tc": 23642880.0, "k_wgrad_tc": 32341760.0}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi style clock / throttle-reason samples during the timed region (NVML)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
            "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10,
            "applications_clocks_setting": 0x2,
        }
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv is not None:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv is not None:
            self.t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_batches(n_batches: int, rank: int):
    """-> list of (points [B*PTS,6] f32, labels [B*PTS] i64) numpy batches of `planes` scenes"""
    from gapartnet_b200 import synthetic

    out = []
    for j in range(n_batches):
        scs = [synthetic.planes(3000 + 100000 * rank + 1000 * j + i, PTS) for i in range(BATCH)]
        out.append((np.concatenate([s.points for s in scs]), np.concatenate([s.sem_labels for s in scs])))
    return out


def algorithmic_bytes_per_scene(counts_per_level, batch):
    """per conv layer fwd+dgrad+wgrad, fp32 feats, int32 [K,n_out] table:
    12*(n_in*Cin + n_out*Cout) + 12*K*n_out (K>1) + 12*K*Cin*Cout (weights counted per step)."""
    D = len(CHANNELS)
    feat = 0.0

    def conv(n_in, n_out, cin, cout, K):
        nonlocal feat
        feat += 12 * (n_in * cin + n_out * cout) + (12 * K * n_out if K > 1 else 0)

    M = counts_per_level
    conv(M[0], M[0], IN_CH, CHANNELS[0], 27)
    for L in range(D):
        c = CHANNELS[L]
        for _ in range(BLOCK_REPEAT):
            conv(M[L], M[L], c, c, 27)
            conv(M[L], M[L], c, c, 27)
        if L + 1 < D:
            c1 = CHANNELS[L + 1]
            conv(M[L], M[L + 1], c, c1, 8)
            conv(M[L + 1], M[L], c1, c, 8)
            conv(M[L], M[L], 2 * c, c, 1)
            conv(M[L], M[L], 2 * c, c, 27)
            conv(M[L], M[L], c, c, 27)
            for _ in range(BLOCK_REPEAT - 1):
                conv(M[L], M[L], c, c, 27)
                conv(M[L], M[L], c, c, 27)
    return feat / batch


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from gapartnet_b200._lib import C
    from gapartnet_b200.engine import SparseUNetEngine
    from gapartnet_b200.network import backbone as mirror
    import gapartnet_b200.spconv.pytorch as sp

    torch.manual_seed(23333)  # gapartnet.yaml:88
    # (the reference's torch.set_float32_matmul_precision('medium'), gapartnet/train.py:6, only governs the torch
    #  Linear heads; measured here it makes the 16->10 head GEMMs slower - 7.04 vs 6.89 ms/step - so the heads
    #  stay on exact-fp32 cuBLAS)
    net = mirror.build_sparse_unet(sp, IN_CH, CHANNELS, BLOCK_REPEAT).to(dev)
    head_w = (torch.randn(NUM_CLASSES, CHANNELS[0], device=dev) * 0.1)
    head_b = torch.zeros(NUM_CLASSES, device=dev)
    head_gw, head_gb = torch.zeros_like(head_w), torch.zeros_like(head_b)
    N = BATCH * PTS
    eng = SparseUNetEngine(net, batch=BATCH, max_points=N, spatial_shape=(SHAPE,) * 3, voxel_size=VOXEL,
                           in_channels=IN_CH)
    off = torch.arange(BATCH + 1, dtype=torch.int64, device=dev) * PTS
    eng.batch_offsets.copy_(off)

    n_rot = 4
    host = make_batches(n_rot, rank)
    pin_pts = [torch.from_numpy(p).pin_memory() for p, _ in host]
    pin_lab = [torch.from_numpy(l).pin_memory() for _, l in host]
    dev_pts = [p.to(dev) for p in pin_pts]
    dev_lab = [l.to(dev) for l in pin_lab]
    labels = torch.empty(N, dtype=torch.int64, device=dev)
    loss_buf = torch.zeros(1, device=dev)
    loss_host = torch.zeros(1).pin_memory()

    def step_body():
        """voxelize + rulebooks + fwd + sem head/CE + bwd (+ allreduce); everything on the stream."""
        eng.flat_grad.zero_()
        eng.build_levels(overlap=True)     # deeper rulebooks + weight packing on a side stream, joined in run_forward
        feat = eng.run_forward()
        # semantic head (network/model.py:104,160-166) + cross-entropy, backward written out by hand
        logits = torch.addmm(head_b, feat, head_w.t())
        logp = torch.log_softmax(logits, dim=1)
        loss_buf.copy_(-(logp.gather(1, labels[:, None]).mean()).reshape(1))
        dlog = torch.softmax(logits, dim=1)
        dlog.scatter_add_(1, labels[:, None], torch.full((N, 1), -1.0, device=dev))
        dlog.mul_(1.0 / N)
        torch.mm(dlog.t(), feat, out=head_gw)
        head_gb.copy_(dlog.sum(0))
        torch.mm(dlog, head_w, out=eng.d_pc_feature)
        eng.run_backward()

    # ---- warm-up eager, then capture the step in a CUDA graph -----------------------------------
    eng.points.copy_(dev_pts[0])
    labels.copy_(dev_lab[0])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step_body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    counts = eng.calibrate()   # per-level row counts -> split-K launch hints for the deep levels
    if args.ncu_step:
        # exactly one eager step between cudaProfilerStart/Stop (ncu --profile-from-start off)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step_body()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        print(json.dumps({"ncu_step": True, "level_rows": counts}))
        return
    use_graph = not args.no_graph
    graph = None
    l0 = C.gp_launch_count()
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_body()
    else:
        step_body()
    launches_per_step = int(C.gp_launch_count() - l0)

    # end-to-end feed: pinned host batches are copied by a copy stream into two staging buffers, one step ahead of
    # the compute stream (every step's H2D happens inside the timed region, like a prefetching DataLoader)
    copy_stream = torch.cuda.Stream()
    stage_pts = [torch.empty_like(dev_pts[0]) for _ in range(2)]
    stage_lab = [torch.empty_like(dev_lab[0]) for _ in range(2)]
    ev_h2d = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def h2d(i):
        k, j = i % 2, i % n_rot
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[k])          # the compute stream consumed this staging buffer
            stage_pts[k].copy_(pin_pts[j], non_blocking=True)
            stage_lab[k].copy_(pin_lab[j], non_blocking=True)
            ev_h2d[k].record(copy_stream)

    def step(i, from_host, last=False):
        j = i % n_rot
        cur = torch.cuda.current_stream()
        if from_host:
            k = i % 2
            cur.wait_event(ev_h2d[k])
            eng.points.copy_(stage_pts[k], non_blocking=True)
            labels.copy_(stage_lab[k], non_blocking=True)
            ev_free[k].record(cur)
            if not last:
                h2d(i + 1)
        else:
            eng.points.copy_(dev_pts[j], non_blocking=True)
            labels.copy_(dev_lab[j], non_blocking=True)
        if graph is not None:
            graph.replay()
        else:
            step_body()
        if world > 1:
            dist.all_reduce(eng.flat_grad)
            dist.all_reduce(head_gw)
        if from_host:
            loss_host.copy_(loss_buf, non_blocking=True)

    def timed(from_host):
        if from_host:
            for k in range(2):
                ev_free[k].record(torch.cuda.current_stream())
            h2d(0)
        for i in range(args.warmup):
            step(i, from_host, last=(i == args.warmup - 1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if from_host:
            copy_stream.wait_event(e0)      # the first timed step's H2D is inside the timed region too
            h2d(args.warmup)
        for i in range(args.steps):
            step(args.warmup + i, from_host, last=(i == args.steps - 1))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms

    with ClockSampler(local) as clk:
        ms_res = timed(False)
    ms_e2e = timed(True)
    loss_val = float(loss_host.item())

    pts_per_step = N * world
    value = pts_per_step * args.steps / (ms_res / 1e3)
    e2e_value = pts_per_step * args.steps / (ms_e2e / 1e3)

    # ---- dominant kernels: the L0 SubMConv3d 16->16 launches (forward/dgrad operator k_conv_tc and the weight
    # gradient k_wgrad_tc), timed one by one with CUDA events on the launching stream -------------------------
    roof = None
    if rank == 0:
        M0 = counts[0]
        x = torch.randn(eng.max_rows[0], 16, device=dev)
        y = torch.empty_like(x)
        dyv = torch.randn_like(x)
        w = net.ublock.encoder_blocks[0].conv1[0].weight
        dw = torch.zeros_like(w)
        ws = torch.empty(int(C.gp_conv_tc_workspace_floats(27, 16, 16)), device=dev)
        st = torch.cuda.current_stream().cuda_stream
        nbr, dn = eng.nbr[0], eng.d_n[0]
        C.gp_conv_tc_fwd(x.data_ptr(), 16, 16, w.data_ptr(), 16, 1, 27 * 16, 0, nbr.data_ptr(), nbr.shape[1], 27,
                         dn.data_ptr(), eng.max_rows[0], y.data_ptr(), 16, 16, 0, None, ws.data_ptr(), M0, st)

        def time_launch(fn):
            evs = []
            for _ in range(3 + 10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            return float(np.mean([a.elapsed_time(b) for a, b in evs[3:]]))

        t_conv = time_launch(lambda: C.gp_conv_tc_run(
            x.data_ptr(), 16, 16, ws.data_ptr(), nbr.data_ptr(), nbr.shape[1], 27, dn.data_ptr(), eng.max_rows[0],
            y.data_ptr(), 16, 16, 0, None, M0, 0, st))
        t_wgrad = time_launch(lambda: C.gp_conv_wgrad_tc(
            x.data_ptr(), 16, 16, dyv.data_ptr(), 16, 16, nbr.data_ptr(), nbr.shape[1], 27, dn.data_ptr(),
            eng.max_rows[0], dw.data_ptr(), 16, 1, 27 * 16, M0, st))
        # algorithmic bytes of one launch (SURVEY 8d per-layer figure / 3 passes): X + Y (or dY) + table + W
        alg = 4.0 * (M0 * 16 + M0 * 16 + 27 * M0 + 27 * 16 * 16)
        peak, peak_src = _peaks()

        def entry(name, ms):
            ach = alg / (ms / 1e3) / 1e9
            return {"bound": "hbm", "kernel": name, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": None, "launch_ms": round(ms, 4),
                    "algorithmic_bytes": alg, "peak_source": peak_src}

        roof = entry("k_conv_tc (L0 SubMConv3d 16->16 forward; the same kernel runs every dgrad)", t_conv)
        roof["traffic"] = TRAFFIC_NCU.get("k_conv_tc")
        roof["other_kernels"] = [entry("k_wgrad_tc (L0 SubMConv3d 16->16 weight gradient)", t_wgrad)]
        roof["other_kernels"][0]["traffic"] = TRAFFIC_NCU.get("k_wgrad_tc")
        # whole-step algorithmic traffic against the step time
        per_scene = algorithmic_bytes_per_scene([c / BATCH for c in counts], 1)
        wbytes = 12.0 * sum(p.numel() for n_, p in net.named_parameters() if p.dim() == 5)
        step_bytes = per_scene * BATCH + wbytes
        roof["step_algorithmic_GB"] = round(step_bytes / 1e9, 4)
        roof["step_frac"] = round(step_bytes / (ms_res / args.steps / 1e3) / 1e9 / peak, 4)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_points_per_sec(steps=2, warmup=0, scenes=2)

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_res / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (planes generator, random-init weights seed 23333)",
        "config": {"workload": "cfg3: sparse U-Net backbone fwd+bwd (+voxelize, 13 rulebooks, sem head/CE)",
                   "points_per_scene": PTS, "batch_per_gpu": BATCH, "voxel": VOXEL, "channels": CHANNELS,
                   "block_repeat": BLOCK_REPEAT, "parallelism": f"dp{world}", "cuda_graph": use_graph,
                   "l2": "per-step working set (activations+tables, >1 GB) exceeds the 126 MB L2; 4 rotating input batches",
                   "level_rows": counts},
        "clocks": clk.summary(),
        "e2e": {"value": round(e2e_value, 1), "unit": "points/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                "h2d_bytes_per_step": int(N * IN_CH * 4 + N * 8), "d2h_bytes_per_step": 4, "loss": loss_val},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def cpu_reference_points_per_sec(steps: int, warmup: int, scenes: int):
    """The reference's CPU path for this workload = the oracle port (spconv/epic_ops are not
    installable here, oracle/__init__.py): numpy voxelize + rulebooks + torch-CPU
    gather-mm-index_add U-Net forward/backward on `scenes` 20k-point scenes per step."""
    from gapartnet_b200 import synthetic
    from gapartnet_b200.network import backbone as mirror
    from oracle import spconv_cpu as osp
    from oracle import voxelize as ovox

    # the port's ops are small index_add/mm calls: beyond ~16 threads torch's intra-op pool only
    # thrashes (measured on the 128-core box: 128 threads were 50x slower than 8), so `cores`
    # reports the threads actually used
    cores = min(os.cpu_count() or 1, 16)
    torch.set_num_threads(cores)
    torch.manual_seed(23333)
    net = mirror.build_sparse_unet(osp, IN_CH, CHANNELS, BLOCK_REPEAT)
    head = torch.nn.Linear(CHANNELS[0], NUM_CLASSES)

    def one(seed):
        scs = [synthetic.planes(seed + i, PTS) for i in range(scenes)]
        t0 = time.perf_counter()
        feats, idx, pcid, off = [], [], [], 0
        shape = [SHAPE] * 3
        for b, sc in enumerate(scs):
            vf, vc, pid, rng = ovox.apply_voxelization(sc.points, [VOXEL] * 3, min_shape=SHAPE)
            feats.append(vf)
            idx.append(np.concatenate([np.full((vc.shape[0], 1), b, np.int32), vc], 1))
            pcid.append(pid + off)
            off += vc.shape[0]
            shape = np.maximum(shape, rng).tolist()
        x = osp.SparseConvTensor(torch.from_numpy(np.concatenate(feats)), torch.from_numpy(np.concatenate(idx)),
                                 shape, scenes)
        net.zero_grad()
        y = net(x).features[torch.from_numpy(np.concatenate(pcid))]
        lab = torch.from_numpy(np.concatenate([s.sem_labels for s in scs]))
        loss = torch.nn.functional.cross_entropy(head(y), lab)
        loss.backward()
        return time.perf_counter() - t0

    for i in range(warmup):
        one(9000 + 10 * i)
    ts = [one(9500 + 10 * i) for i in range(steps)]
    t = float(np.sum(ts))
    return {"value": round(scenes * PTS * steps / t, 1), "unit": "points/s", "cores": cores, "kind": "port",
            "sample": f"{steps} step(s) x {scenes} scene(s) of {PTS} pts (same graph, voxel {VOXEL}), "
                      f"{t / steps:.2f} s/step, torch {torch.get_num_threads()} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    scenes = 2
    t0 = time.perf_counter()
    r = cpu_reference_points_per_sec(steps=args.steps, warmup=args.warmup, scenes=scenes)
    dt = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * scenes * PTS / r["value"], 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (planes generator)",
        "config": {"workload": "cfg3: sparse U-Net backbone fwd+bwd (+voxelize, rulebooks, sem head/CE) - CPU oracle port "
                               "(spconv/epic_ops are not vendored/installable: oracle/__init__.py)",
                   "points_per_scene": PTS, "scenes_per_step": scenes, "voxel": VOXEL, "channels": CHANNELS,
                   "block_repeat": BLOCK_REPEAT},
        "cpu_baseline": r,
        "e2e": {"value": r["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(dt, 1),
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-step", action="store_true", help="run one eager step inside cudaProfilerStart/Stop and exit")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
