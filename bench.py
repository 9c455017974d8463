#!/usr/bin/env python
"""bench.py — MinkUNet34C (Res16UNet34C) fwd+bwd+SGD voxels/s on synthetic ScanNet-shaped plenoxel
scenes (BASELINE.json configs[1]; metric "MinkUNet fwd+bwd voxels/sec").

  python bench.py --gpus N --steps K --warmup W            our arm (sm_100a kernels behind the ME surface)
  python bench.py --impl reference ...                      CPU arm: the oracle port of ME's CPU algorithm
                                                            (ME itself is not installable here)

A step = one pass of the hot path over one batch: voxel quantisation + hashing, kernel maps, every
conv fwd/dgrad/wgrad, BN/ReLU/pooling, loss, gradient all-reduce (N>1) and the SGD update.  The
coordinate manager is rebuilt every step, as in the reference where each step makes a new TensorField.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "MinkUNet34C fwd+bwd voxels/sec"
UNIT = "voxels/s"

# The stated parity bar of each operand precision and where it is proven; the whole-network
# figures are Res16UNet34C at 2 x 200 K voxels against the fp32 CUDA-core path on the same device
# (tests/test_gpu_scale.py, profiles/r2_precision_at_scale.md).
PARITY_NOTE = {
    "bf16": {"per_layer_bar": "|d| <= 3e-3 * max|ref| (fwd, dgrad, wgrad; measured 2.2e-3 .. 2.7e-3)",
             "test": "tests/test_gpu_parity.py::test_conv_bf16_tensor_core, tests/test_gpu_scale.py",
             "whole_network_at_scale": "logits cos 0.9993, all-parameter gradient cos 0.903 (tf32: 0.99999 / 0.985)",
             "graph": "conv + BatchNorm as one autograd node, hollow rows, re-computed ReLU masks, epilogue statistics, "
                      "symmetric dgrad, residual gradients in dgrad's epilogue (ops.<knob>; held to the plain graph in "
                      "tests/test_gpu_fusion.py)"},
    "tf32": {"per_layer_bar": "|d| <= 3e-3 * max|ref| (measured 7.4e-4 .. 8.5e-4)",
             "test": "tests/test_gpu_parity.py::test_conv_tf32_tensor_core, tests/test_gpu_scale.py",
             "whole_network_at_scale": "logits cos 0.99999, all-parameter gradient cos 0.985"},
    "fp32": {"per_layer_bar": "|d| <= 1e-4 * (1 + |ref|)", "test": "tests/test_gpu_parity.py::test_conv_fp32_vs_fp64_oracle"},
}


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    # TF32: not in MEASURED_PEAKS.json; measured with the same protocol (torch.matmul 8192^3, allow_tf32) on this
    # pool's B200 in round 2 (scripts/measure_tf32_peak.py -> profiles/r2_tf32_peak.json)
    t = ROOT / "profiles" / "r2_tf32_peak.json"
    tf32 = json.loads(t.read_text()) if t.exists() else {"tf32_tflops": 749.1, "tf32_tflops_sustained": 592.1}
    out = {"tf32_burst": tf32["tf32_tflops"], "tf32_sustained": tf32["tf32_tflops_sustained"]}
    if p.exists():
        d = json.loads(p.read_text())
        out.update({"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                    "source": "measured"})
    else:
        out.update({"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"})
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]  # upper half ~ samples under load
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arm: oracle port of ME's CPU algorithm (gather -> sgemm -> scatter, OpenMP kernel maps)
# ---------------------------------------------------------------------------
def cpu_step_fn(voxels: int, seed: int = 777, forward_only: bool = False):
    import numpy as np
    import torch
    from nerf_downstream_b200 import models, synth
    from oracle import nets

    torch.manual_seed(0)
    coords, feats, labels = synth.room_batch(seed, 1, voxels)
    model = models.Res16UNet34C(27, 20)
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k)
              for k, v in model.state_dict().items()}
    f = torch.from_numpy(feats)
    y = torch.from_numpy(labels)

    def step():
        if forward_only:   # the full 1 M-voxel scene on the CPU is timed forward-only (labelled)
            with torch.no_grad():
                logits = nets.resunet_forward(params, coords, f, use_c=True)
                return float(torch.nn.functional.cross_entropy(logits, y, ignore_index=255))
        for p in params.values():
            p.grad = None
        logits = nets.resunet_forward(params, coords, f, use_c=True)
        loss = torch.nn.functional.cross_entropy(logits, y, ignore_index=255)
        loss.backward()
        return float(loss.detach())

    return step, coords.shape[0]


def cpu_baseline(voxels: int, budget_s: float = 20.0):
    import torch
    step, n = cpu_step_fn(voxels)
    step()  # warm-up (builds the C oracle, touches MKL)
    t0 = time.perf_counter()
    reps = 0
    while True:
        step()
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 5:
            break
    dt = (time.perf_counter() - t0) / reps
    return {"value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"Res16UNet34C(27,20) fwd+bwd on one {n}-voxel synthetic room, fp32, {reps} reps; "
                      "CPU restatement of ME's algorithm (ME not installable)",
            "host_cpus": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's CPU path = oracle port, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    voxels = args.cpu_voxels
    step, n = cpu_step_fn(voxels, forward_only=args.cpu_forward_only)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    what = "FORWARD ONLY (fwd+bwd is ~3x; SURVEY 8d prescribes forward-only x3 for the full scene)" if args.cpu_forward_only \
        else "fwd+bwd"
    same = n >= 0.95 * args.voxels
    sample = (f"each step = Res16UNet34C(27,20) {what} on one {n}-voxel synthetic room "
              + ("(the workload's own scene size)" if same else f"(bounded sample of the {args.voxels}-voxel workload)")
              + ", fp32, oracle port of ME's CPU algorithm (ME not installable)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"MinkUNet34C (Res16UNet34C 27->20) fwd+bwd, synthetic ScanNet-shaped plenoxel "
                                   f"scenes, {args.voxels} voxels/scene, {args.scenes} scene(s)/GPU",
                       "sample_voxels": n, "same_config": bool(same and not args.cpu_forward_only),
                       "same_scene_size": bool(same), "forward_only": bool(args.cpu_forward_only)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import ctypes

    import numpy as np
    import torch
    import torch.distributed as dist

    from nerf_downstream_b200 import lib as L
    from nerf_downstream_b200 import me as ME
    from nerf_downstream_b200 import models, ops, pipeline, synth, trainer

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    lib = L.load()
    if args.sort_rows:
        ops.sort_rows = True

    torch.manual_seed(0)
    model = models.Res16UNet34C(27, 20).to(dev).train()
    tr = trainer.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4)

    if args.geometry == "faithful":   # the reference's real ScanNet-plenoxel coordinate transform
        coords, feats, labels = synth.faithful_room_batch(777 + rank, args.scenes, args.voxels, args.scene_scale)
    else:
        coords, feats, labels = synth.room_batch(777 + rank, args.scenes, args.voxels, shuffle=args.shuffle)
    h_coords = torch.from_numpy(coords).pin_memory()
    h_feats = torch.from_numpy(feats).pin_memory()
    h_labels = torch.from_numpy(labels).pin_memory()
    d_coords, d_feats, d_labels = h_coords.to(dev), h_feats.to(dev), h_labels.to(dev)
    voxels_per_step = [0]
    # end-to-end input: the batch as PeRFception stores it (int32 links + uint8 SH + uint8 labels, 32 B / voxel),
    # decoded on the device into the collated tensors (spc_plenoxel_decode) — --e2e-input float: the float tensors
    compact = None
    if args.e2e_input == "compact" and args.geometry == "dense":
        recs = synth.compact_records(coords, feats, labels)
        compact = {"recs": [(x[0], x[4], len(x[1])) for x in recs],   # (batch index, grid, rows)
                   "links": torch.from_numpy(np.concatenate([r[1] for r in recs])).pin_memory(),
                   "sh": torch.from_numpy(np.concatenate([r[2] for r in recs])).pin_memory(),
                   "labels": torch.from_numpy(np.concatenate([r[3] for r in recs])).pin_memory()}

    def step(c, f, y):
        field = ME.TensorField(coordinates=c, features=f)
        if args.fused_head:     # slice + loss (+ gradient of the slice) as one kernel over the points (spc_seg_head_fwd)
            loss = pipeline.seg_head_loss(model.forward_sparse(field), field, y, 255)
        else:
            logits = model(field)
            loss = ops.cross_entropy(logits, y, ignore_index=255)
        tr.backward_and_step(loss)
        key = field.coordinate_manager.get_unique_coordinate_map_key(1)
        voxels_per_step[0] = field.coordinate_manager.size(key)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def path_counts(reset):
        buf = (ctypes.c_longlong * 3)()
        lib.spc_conv_path_counts(ctypes.cast(buf, ctypes.c_void_p), int(reset))
        return [int(v) for v in buf]

    peaks = _peaks()
    copy_stream = torch.cuda.Stream(device=dev)

    def measure(precision: str, with_clocks: bool):
        """Warm-up, device-resident timed region, end-to-end timed region and the per-kernel profile of ONE
        operand precision.  Returns the fields of the JSON line that depend on it."""
        ops.set_default_precision(precision)
        # ---- warm-up, then the device-resident timed region -------------------------------------
        for _ in range(max(args.warmup, 3)):
            step(d_coords, d_feats, d_labels)
        sampler = ClockSampler(local_rank)
        if rank == 0 and with_clocks:
            sampler.start()
        launches0 = L.launch_count()
        path_counts(True)
        host_s = [0.0]

        def host_timed_step():
            t0 = time.perf_counter()
            step(d_coords, d_feats, d_labels)
            host_s[0] += time.perf_counter() - t0  # host time to ISSUE the step (it blocks only in the map-size read-back)

        ms = timed(args.steps, host_timed_step)
        launches = L.launch_count() - launches0
        # host time to ISSUE one step with an idle device in front of it: inside the timed region the map-size
        # read-backs make the host wait for the previous step's kernels, so `host_issue_ms_per_step` there tracks the
        # device time; this one is what the Python layer itself costs per step (median of 3)
        iso = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step(d_coords, d_feats, d_labels)
            iso.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
        host_isolated_ms = 1e3 * sorted(iso)[1]
        routes = path_counts(False)
        clocks = sampler.stop() if (rank == 0 and with_clocks) else None
        vox = torch.tensor([float(voxels_per_step[0])], device=dev)
        if world > 1:
            dist.all_reduce(vox, op=dist.ReduceOp.SUM)
        total_voxels = float(vox.item())
        value = total_voxels * args.steps / (ms * 1e-3)

        # ---- end to end through the public API with HOST buffers ------------------------------------
        # Every step's inputs travel host -> device from pinned memory inside the timed region and the
        # loss is read back every step.  As a training input pipeline does (pin_memory + non_blocking),
        # the copy of batch i+1 is enqueued on a copy stream while step i computes (two device buffers);
        # batch i+1 is never touched before its copy event, and buffer reuse is safe because loss.item()
        # of step i-1 has synchronised the device.
        host_src = (compact["links"], compact["sh"], compact["labels"]) if compact else (h_coords, h_feats, h_labels)
        bufs = [tuple(torch.empty_like(t, device=dev) for t in host_src) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        turn = [0]
        dec_c = torch.empty_like(d_coords) if compact else None
        dec_f = torch.empty_like(d_feats) if compact else None

        def enqueue_copy(slot):
            with torch.cuda.stream(copy_stream):
                for d, h in zip(bufs[slot], host_src):
                    d.copy_(h, non_blocking=True)
                ready[slot].record(copy_stream)

        def e2e_step():
            slot = turn[0] & 1
            turn[0] += 1
            torch.cuda.current_stream().wait_event(ready[slot])
            enqueue_copy(slot ^ 1)  # next step's batch, overlapped with this step
            if compact:
                links, sh, lab8 = bufs[slot]
                row = 0
                for b, reso, n in compact["recs"]:   # each record decodes into its rows of the collated batch
                    pipeline.plenoxel_decode(links[row:row + n], sh[row:row + n], 2.0 / 255.0, -1.0, reso, batch_index=b,
                                             affine=(1, 0, 0, 0, 1, 0, 0, 0, 1, 0.5, 0.5, 0.5),
                                             out_coords=dec_c[row:row + n], out_feats=dec_f[row:row + n])
                    row += n
                c, f, y = dec_c, dec_f, lab8.long()
            else:
                c, f, y = bufs[slot]
            loss = step(c, f, y)
            return loss.item()  # device -> host read of the step's result

        enqueue_copy(0)
        e2e_step()
        e2e_steps = max(3, args.steps // 2)
        ms_e2e = timed(e2e_steps, e2e_step)
        torch.cuda.current_stream().wait_stream(copy_stream)
        e2e_value = total_voxels * e2e_steps / (ms_e2e * 1e-3)
        h2d = sum(t.numel() * t.element_size() for t in host_src)
        d2h = 4 + 8 * 5  # loss scalar + the per-level map sizes the host reads back (2 int32 each)

        # ---- per-kernel-class device times inside a (separately) timed region -> roofline -------------
        roofline = kmap_ms = others = None
        classes = {}
        prof = ops.KernelProfiler() if rank == 0 else None
        ops.set_profiler(prof)
        for _ in range(2):  # every rank steps (the gradient all-reduce is collective); rank 0 records
            step(d_coords, d_feats, d_labels)
        torch.cuda.synchronize()
        ops.set_profiler(None)
        if rank == 0:
            classes = prof.summary()
            det = prof.summary_detail()
            if args.detail:
                for (name, detail), v in sorted(det.items(), key=lambda kv: -kv[1]["ms"])[:40]:
                    t = v["ms"] / v["n"]
                    print(f"# [{precision}] {name:12s} {detail:34s} n={v['n']:3d} avg={t:8.3f} ms  "
                          f"{v['flops'] / v['n'] / (t * 1e-3) / 1e12 if t > 0 else 0:7.1f} TFLOP/s  "
                          f"{v['bytes'] / v['n'] / (t * 1e-3) / 1e9 if t > 0 else 0:7.0f} GB/s(alg)", file=sys.stderr)
            tensor_peak = peaks["bf16_sustained"] if precision == "bf16" else peaks["tf32_sustained"]
            tensor_note = (f"{peaks['source']} sustained cuBLAS bf16 rate (kernel timed inside a long step)" if precision == "bf16"
                           else "sustained cuBLAS TF32 rate measured on this pool with the MEASURED_PEAKS.json protocol "
                                "(profiles/r2_tf32_peak.json)")
            # BASELINE.json's second metric: 3^3 stride-1 kernel map at tensor stride 1 over the full scene
            km = [(d, v) for (n, d), v in det.items() if n == "kernel_map" and d.startswith("K27 ") and d.endswith("ts1")]
            if km:
                d, v = max(km, key=lambda kv: int(kv[0].split(" M")[1].split()[0]))
                kmap_ms = {"value": v["ms"] / v["n"], "unit": "ms", "map": "3^3 stride 1 @ tensor stride 1",
                           "voxels": int(d.split(" M")[1].split()[0]),
                           "algorithmic_gbs": v["bytes"] / v["n"] / (v["ms"] / v["n"] * 1e-3) / 1e9}
            step_ms_prof = sum(v["ms"] for v in classes.values()) / 2
            if det:
                # dominant kernel = the kernel class with the largest device time in the step, reported on the
                # layer shape that takes most of that time
                top_class = max(classes.items(), key=lambda kv: kv[1]["ms"])[0]
                (name, detail), s_ = max(((k, v) for k, v in det.items() if k[0] == top_class), key=lambda kv: kv[1]["ms"])
                t = s_["ms"] * 1e-3
                shape = detail.split(" P")[0]
                traffic = None
                tfile = ROOT / "profiles" / "roofline_traffic.json"
                if tfile.exists():  # ncu DRAM bytes per row of this kernel shape x rows of the reported launch
                    ent = json.loads(tfile.read_text()).get(f"{precision} {name} {shape.split(' M')[0]}")
                    if ent and " M" in shape:
                        traffic = ent["dram_bytes_per_row"] * int(shape.split(" M")[1].split()[0])
                common = {"kernel": f"{name} {shape}".strip(), "traffic": traffic, "launches": s_["n"],
                          "avg_launch_ms": s_["ms"] / s_["n"], "share_of_step": s_["ms"] / 2 / step_ms_prof,
                          "class_share_of_step": classes[name]["ms"] / 2 / step_ms_prof}
                if name.startswith("conv") and s_["flops"] > 0:
                    ach = s_["flops"] / t / 1e12
                    roofline = {**common, "bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s",
                                "frac": ach / tensor_peak, "peak_note": tensor_note, "algorithmic_gbs": s_["bytes"] / t / 1e9}
                else:
                    ach = s_["bytes"] / t / 1e9
                    roofline = {**common, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": ach / peaks["hbm_gbs"], "peak_note": peaks["source"]}
                # context: the five largest (kernel, shape) entries with their own roofline fractions
                others = []
                for (n_, d_), v in sorted(det.items(), key=lambda kv: -kv[1]["ms"])[:5]:
                    t_ = v["ms"] * 1e-3
                    if n_.startswith("conv") and v["flops"] > 0:
                        others.append({"kernel": f"{n_} {d_.split(' P')[0]}", "ms_per_step": round(v["ms"] / 2, 3),
                                       "tflops": round(v["flops"] / t_ / 1e12, 1),
                                       "frac": round(v["flops"] / t_ / 1e12 / tensor_peak, 3)})
                    else:
                        others.append({"kernel": f"{n_} {d_}".strip(), "ms_per_step": round(v["ms"] / 2, 3),
                                       "gbs": round(v["bytes"] / t_ / 1e9, 0),
                                       "frac": round(v["bytes"] / t_ / 1e9 / peaks["hbm_gbs"], 3)})
        return {"value": value, "ms_per_step": ms / args.steps, "total_voxels": total_voxels, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / e2e_steps,
                        "input": ("plenoxel records as stored (int32 links + uint8 SH + uint8 labels), decoded on the device "
                                  "by spc_plenoxel_decode inside the timed region") if compact else
                                 "collated float32 coordinates + float32 features + int64 labels"},
                "gpu_launches": int(launches), "host_issue_ms_per_step": 1e3 * host_s[0] / args.steps,
                "host_issue_ms_isolated_step": host_isolated_ms,
                # convolution launches per step by route: anything under "cuda_core_fp32" is a shape the tensor-core
                # kernels do not take (ops.SparseConvFn / conv_api.cu routing), "tf32" under a bf16 run a precision
                # fallback — both would be silent otherwise
                "conv_routes_per_step": {"tcgen05_bf16": routes[0] / args.steps, "tcgen05_tf32": routes[1] / args.steps,
                                         "cuda_core_fp32": routes[2] / args.steps},
                "roofline": roofline, "kernel_map_build_ms": kmap_ms, "other_kernels": others,
                "kernel_classes_ms_per_step": {k: round(v["ms"] / 2, 4) for k, v in
                                               sorted(classes.items(), key=lambda kv: -kv[1]["ms"])}}

    main_res = measure(args.precision, True)
    alt_res = None
    alt = {"bf16": "tf32", "tf32": "bf16"}.get(args.precision)
    if alt is not None and not args.no_alt_precision:
        alt_res = measure(alt, False)

    if args.host_profile and rank == 0:
        import cProfile
        import pstats
        ops.set_default_precision(args.precision)
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(3):
            step(d_coords, d_feats, d_labels)
        pr.disable()
        torch.cuda.synchronize()
        st = pstats.Stats(pr, stream=sys.stderr)
        st.sort_stats("tottime").print_stats(45)
        st.sort_stats("cumulative").print_stats(45)

    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu = cpu_baseline(args.cpu_voxels) if (world == 1 and not args.no_cpu_baseline) else None
        dt = {"tf32": "tf32", "bf16": "bf16", "fp32": "f32"}
        line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": dt[args.precision],
                "data": "synthetic",
                "config": {"workload": f"MinkUNet34C (Res16UNet34C 27->20) fwd+bwd+SGD, synthetic ScanNet-shaped "
                                       f"plenoxel scenes, {args.voxels} voxels/scene, {args.scenes} scene(s)/GPU, "
                                       f"BASELINE.json configs[1]"
                                       + (f" — config 2B geometry, scene_scale {args.scene_scale}"
                                          if args.geometry == "faithful" else ""),
                           "voxels_per_step_all_gpus": main_res["total_voxels"], "parallelism": f"dp{world}",
                           "l2_policy": "inputs_exceed_l2 (activations per step >> 126 MB)",
                           "voxel_order": "shuffled" if args.shuffle else "raster (as the reference loaders deliver)",
                           "engine_row_order": ("on: " + str(ops.sort_stats)) if ops.sort_rows else "off (first occurrence, as ME's CPU maps)",
                           "precision": args.precision},
                # which bar the operand precision of `value` meets (tensor-core modes within
                # 3e-3 * max|ref| per layer) and the test that proves it; whole-network agreement at realistic scale
                "parity": PARITY_NOTE.get(args.precision),
                "clocks": main_res["clocks"],
                "e2e": main_res["e2e"],
                "gpu_launches": main_res["gpu_launches"],
                "host_issue_ms_per_step": main_res["host_issue_ms_per_step"],
                "host_issue_ms_isolated_step": main_res["host_issue_ms_isolated_step"],
                "conv_routes_per_step": main_res["conv_routes_per_step"],
                "roofline": main_res["roofline"],
                "kernel_map_build_ms": main_res["kernel_map_build_ms"],
                "other_kernels": main_res["other_kernels"],
                "cpu_baseline": cpu,
                "kernel_classes_ms_per_step": main_res["kernel_classes_ms_per_step"]}
        if alt_res is not None:
            line["alt_precision"] = {"dtype": dt[alt], "value": alt_res["value"], "unit": UNIT,
                                     "ms_per_step": alt_res["ms_per_step"], "e2e": alt_res["e2e"],
                                     "roofline": alt_res["roofline"], "parity": PARITY_NOTE.get(alt),
                                     "conv_routes_per_step": alt_res["conv_routes_per_step"],
                                     "gpu_launches": alt_res["gpu_launches"],
                                     "kernel_classes_ms_per_step": alt_res["kernel_classes_ms_per_step"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--voxels", type=int, default=1_000_000, help="voxels per scene")
    ap.add_argument("--scenes", type=int, default=2,
                    help="scenes per GPU per step (SURVEY.md §8d config 2: B = 1 and B = 2; the reference trains at 8)")
    ap.add_argument("--precision", default="bf16", choices=["tf32", "bf16", "fp32"],
                    help="conv operand precision: bf16 (default; fp32 accumulate), tf32, or fp32 CUDA cores")
    ap.add_argument("--cpu-voxels", type=int, default=20_000, help="scene size of the bounded CPU sample")
    ap.add_argument("--cpu-forward-only", action="store_true",
                    help="--impl reference: time the forward pass only (with --cpu-voxels 1000000: the workload's own "
                         "scene size on the CPU, SURVEY 8d)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-input", default="compact", choices=["compact", "float"],
                    help="what the end-to-end leg copies host -> device each step: the plenoxel records as stored on disk "
                         "(links + u8 SH + u8 labels, decoded on the device) or the collated float tensors")
    ap.add_argument("--no-alt-precision", action="store_true",
                    help="skip the second measurement in the other tensor-core operand precision (alt_precision)")
    ap.add_argument("--detail", action="store_true", help="per-layer kernel times on stderr")
    ap.add_argument("--host-profile", action="store_true", help="cProfile of 3 steps on stderr (host overhead)")
    ap.add_argument("--geometry", default="dense", choices=["dense", "faithful"],
                    help="dense = config 2A (headline: dense 2 cm surface); faithful = config 2B (256^3 plenoxel grid, even "
                         "lattice, (c/256*2-1)/scene_scale/0.02: samples ~2.3 voxels apart, centre-tap-only maps at stride 1)")
    ap.add_argument("--scene-scale", type=float, default=0.34, help="scene_scale of --geometry faithful (median 0.34)")
    ap.add_argument("--fused-head", action="store_true",
                    help="loss through the fused segmentation head (forward_sparse + spc_seg_head_fwd) instead of "
                         "slice -> cross-entropy; off by default until re-measured inside the step")
    ap.add_argument("--sort-rows", action="store_true",
                    help="engine-side row order (ops.sort_rows): rows of large thin maps grouped by neighbour mask; pays on "
                         "--geometry faithful, never triggers on the dense default workload")
    ap.add_argument("--shuffle", action="store_true",
                    help="deliver voxels in random order instead of the loaders' raster order (adversarial locality)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
