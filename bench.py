#!/usr/bin/env python
"""Benchmark of the B200-native relation-classification path (TEMPURA PredCLS fwd+bwd).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

Metric (BASELINE.json): relation pairs / second of a full training step: PredCLS forward + loss + backward
(+ the gradient all-reduce when N > 1) + gradient clipping + AdamW update.  Workload per GPU = BASELINE.json configs[1]: a batch of 64 synthetic
Action-Genome-shaped videos (32 frames, 6-10 pairs/frame, ~16.4 k pairs); N GPUs each take 64 videos
(weak scaling; N = 8 is configs[3], 512 videos per step).  One JSON line is printed by rank 0.

  value     pairs/s with the step's inputs already resident in HBM (CUDA events, max over ranks)
  e2e       pairs/s through the public module API with HOST inputs: every step copies the batch from
            pinned host memory (double-buffered on a copy stream), runs forward + loss + backward and
            reads the loss back to the host
  roofline  the dominant kernel (tcgen05 GEMM): algorithmic 2MNK FLOPs of every GEMM launch in the
            timed region / CUDA-event duration of those launches, vs MEASURED_PEAKS.json
  cpu_baseline  oracle/ (CPU restatement of the reference, pinned to it by golden vectors) timed on
            the box's host cores on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "relation pairs/sec (PredCLS fwd+bwd)"
UNIT = "pairs/s"
MODEL_KW = dict(mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                enc_layer_num=1, dec_layer_num=3, obj_mem_compute=False, rel_mem_compute="joint",
                mem_fusion="late", selection="manual", selection_lambda=0.5, take_obj_mem_feat=False,
                obj_head="gmm", rel_head="gmm", K=6, tracking=False)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--videos", type=int, default=64, help="videos per GPU per step")
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--cpu-sample-videos", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lib-sample-videos", type=int, default=16, help="videos of the PyTorch-eager library baseline")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-teatgt", action="store_true", help="leave out the TEAT-GT (TokenGT) section of the line")
    ap.add_argument("--longclip", action="store_true", help="also run BASELINE configs[4] (always run at 8 GPUs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-consistency", action="store_true", help="leave out the temporal-consistency regulariser")
    ap.add_argument("--profile", action="store_true", help="1 warm-up + --steps steps, no e2e/cpu (for ncu runs)")
    ap.add_argument("--gemm-log", default=None, help="write the per-shape GEMM timing table to this file")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle restatement) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(video_indices, frames, steps, warmup, seed=1123):
    """pairs/s of oracle fwd + loss + bwd, one video per forward like the reference trainer
    (TEMPURA_train.py:152-225), train mode (dropout + GMM noise on), all host threads."""
    import torch
    from b200vsgg import synthetic
    from oracle.tempura_oracle import TempuraOracle, tempura_losses
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = TempuraOracle(obj_classes=synthetic.ag_object_classes(), dropout=0.1, **MODEL_KW)
    synthetic.seeded_init_(model, seed)
    model.train()
    entries = [synthetic.make_video_entry(i, frames, (6, 10)) for i in video_indices]
    labels = [synthetic.build_gt_tensors(e) for e in entries]
    pairs = sum(e["pair_idx"].shape[0] for e in entries)

    def one_pass():
        for e, (att, spa, con) in zip(entries, labels):
            model.zero_grad(set_to_none=True)
            pred = model(dict(e), phase="train")
            loss = sum(tempura_losses(pred, att, spa, con).values())
            loss.backward()

    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = time.perf_counter() - t0
    return pairs * steps / dt, dt / steps, pairs, cores, torch.get_num_threads()


def library_baseline_rates(video_indices, frames, dev, seed=1123, passes=2):
    """The "library baseline to beat" of SURVEY.md §2a / §8(d): the SAME modules as the CPU arm (oracle/, the
    vectorised restatement of the reference — no per-frame Python loops or host syncs, so it is a stronger opponent
    than the reference's own GPU code) moved to the B200 and run by PyTorch eager through cuBLAS / cuDNN / ATen:
    fp32 with TF32 tensor cores, and bf16 autocast.  One video per forward + loss + backward like the reference
    trainer (TEMPURA_train.py:152-225), gradients accumulated, then clip_grad_norm_(5) + torch.optim.AdamW once per
    pass (generous: the reference steps its Python-loop AdamW after every video; no consistency regulariser — the
    reference's TEMPURA has none).  CUDA events, outside the headline timed region.  Checker code, never product."""
    import torch
    import torch.nn.functional as F
    from b200vsgg import synthetic
    from oracle.tempura_oracle import TempuraOracle
    model = TempuraOracle(obj_classes=synthetic.ag_object_classes(), dropout=0.1, **MODEL_KW)
    synthetic.seeded_init_(model, seed)
    model = model.to(dev).train()
    for p in model.object_classifier.parameters():
        p.requires_grad_(False)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=0.1)
    entries = [synthetic.make_video_entry(i, frames, (6, 10), device=dev, big_on_device=dev) for i in video_indices]
    labels = [synthetic.build_gt_tensors(e, dev) for e in entries]
    pairs = sum(e["pair_idx"].shape[0] for e in entries)

    def one_pass(autocast):
        opt.zero_grad(set_to_none=True)
        for e, (att, spa, con) in zip(entries, labels):
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                pred = model(dict(e), phase="train")
            loss = (F.cross_entropy(pred["attention_distribution"].float(), att)
                    + F.binary_cross_entropy(pred["spatial_distribution"].float(), spa)
                    + F.binary_cross_entropy(pred["contacting_distribution"].float(), con))
            loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 5.0)
        opt.step()

    out = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    try:
        for name, autocast in (("eager_fp32_tf32", False), ("eager_bf16_autocast", True)):
            one_pass(autocast)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(passes):
                one_pass(autocast)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / passes
            out[name] = {"value": pairs / (ms * 1e-3), "unit": UNIT, "ms_per_pass": ms}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out["sample"] = ("%d of the step's videos (%d pairs), 1 warm-up + %d timed passes per precision; oracle/tempura_oracle.py "
                     "modules on cuda (PyTorch %s eager: cuBLAS/cuDNN/ATen), one video per forward+loss+backward, "
                     "one clip+AdamW per pass, no regulariser" % (len(entries), pairs, passes, torch.__version__))
    del model, opt, entries, labels
    torch.cuda.empty_cache()
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vids = list(range(args.cpu_sample_videos))
    rate, sec_per_step, pairs, cores, threads = cpu_reference_rate(vids, args.frames, args.steps, min(args.warmup, 1))
    sample = ("%d of the %d videos/GPU of the workload (%d pairs) per step, oracle/tempura_oracle.py fwd+loss+bwd, "
              "fp32, train mode, one video per forward" % (len(vids), args.videos, pairs))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cores": cores},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# TEAT-GT (TokenGT) arm: BASELINE configs[2] (SGCls fwd+bwd) and the PredCLS fwd+bwd+regulariser step; configs[4]
# (long-clip inference, videos sharded over the ranks) with --longclip or at 8 GPUs
# ------------------------------------------------------------------------------------------------
TEAT_ARGS = dict(num_atoms=1168, num_edges=1, num_output=26, lap_node_id=True, lap_node_id_k=50,
                 lap_node_id_sign_flip=False, lap_node_id_eig_dropout=0.2, rand_node_id=False, rand_node_id_dim=50,
                 orf_node_id=False, orf_node_id_dim=50, type_id=True, encoder_embed_dim=768, encoder_layers=12,
                 encoder_attention_heads=32, encoder_ffn_embed_dim=768, return_attention=True)


def _teat_batch(video_indices, frames, ppf, dev, sgcls):
    import numpy as np
    import torch
    from b200vsgg import objbranch, synthetic, tempura
    entries = []
    for i in video_indices:
        e = synthetic.make_video_entry(i, frames, ppf)
        e.pop("union_feat"), e.pop("spatial_masks")          # TEAT-GT never reads them (lib/teatgt.py:98-141)
        e = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in e.items()}
        if sgcls:
            synthetic.add_sgcls_inputs(e, i)
            objbranch.get_sequence(e, None, None, "sgcls")
        entries.append(e)
    gts = [synthetic.build_gt_tensors(e, dev) for e in entries]
    batch = tempura.collate_entries(entries)
    for k in ("attention_gt", "spatial_gt", "contacting_gt"):
        batch.pop(k, None)
    batch["frame_counts_host"] = torch.bincount(batch["im_idx"].long()).cpu().numpy()
    batch["pair_idx_host"] = batch["pair_idx"].cpu().numpy()
    batch["box_frames_host"] = batch["boxes"][:, 0].cpu().numpy()
    return batch, tuple(torch.cat([g[i] for g in gts]) for i in range(3))


def teatgt_cpu_rate(video_indices, frames, sgcls, seed=1123):
    """pairs/s of the TEAT-GT oracle (oracle/teatgt_oracle.py: the reference's per-clip Python loops, fp32 torch on the
    host cores) for forward + losses + backward, one video per forward like TEATGT_train.py:140-190."""
    import types
    import torch
    from b200vsgg import synthetic
    from oracle.teatgt_oracle import TeatgtOracle, teatgt_losses
    from oracle.tempura_oracle import get_sequence, object_loss
    torch.set_num_threads(os.cpu_count() or 1)
    targs = dict(TEAT_ARGS)
    if sgcls:
        targs.update(encoder_layers=6, encoder_attention_heads=16)
    o = TeatgtOracle(mode="sgcls" if sgcls else "predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                     obj_classes=synthetic.ag_object_classes(), tracking=sgcls, args=types.SimpleNamespace(**targs),
                     with_regulariser=not sgcls)
    synthetic.teatgt_seeded_init_(o, seed)
    o.train()
    pairs, t = 0, 0.0
    for n, i in enumerate([video_indices[0]] + list(video_indices)):      # first pass = warm-up
        e = synthetic.make_video_entry(i, frames, (6, 10))
        e.pop("union_feat"), e.pop("spatial_masks")
        if sgcls:
            synthetic.add_sgcls_inputs(e, i)
            get_sequence(e, "sgcls")
        att, spa, con = synthetic.build_gt_tensors(e)
        t0 = time.perf_counter()
        o.zero_grad(set_to_none=True)
        pred = o(dict(e), phase="train")
        loss = sum(teatgt_losses(pred, att, spa, con).values())
        if sgcls:
            loss = loss + object_loss(pred)
        loss.backward()
        if n > 0:
            t += time.perf_counter() - t0
            pairs += e["pair_idx"].shape[0]
    return pairs / t, pairs, torch.get_num_threads()


def teatgt_section(args, dev, rank, world, peaks):
    """One dict for the bench line: 'sgcls' = BASELINE configs[2] (TEAT-GT SGCls fwd+bwd: object branch + 6-layer /
    16-head TokenGT + regulariser), 'predcls' = TEAT-GT PredCLS fwd+bwd with the consistency regulariser (12 layers / 32
    heads), each a FULL training step (losses, backward, gradient all-reduce when N > 1, clip + AdamW) on 64 videos per
    GPU; 'longclip' = configs[4] (inference, 8 videos x 256 frames x 32 pairs per GPU, no collective)."""
    import types
    import numpy as np
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from b200vsgg import ddp, ops, synthetic, teatgt
    from b200vsgg.optim import FusedAdamW

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())

    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    out = {"note": "same rules as the headline: CUDA events, max over ranks, whole-job pairs/s, synthetic AG-shaped videos, "
                   "inputs resident in HBM; roofline fractions vs %.0f TFLOP/s (%s)" % (
                       peak_tf, "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback")}

    def run_train(mode):
        sgcls = mode == "sgcls"
        targs = dict(TEAT_ARGS)
        if sgcls:
            targs.update(encoder_layers=6, encoder_attention_heads=16)     # teatgt_config.py:11-14
        m = teatgt.TEAT_GT(mode=mode, attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                           obj_classes=synthetic.ag_object_classes(), tracking=sgcls, args=types.SimpleNamespace(**targs))
        synthetic.teatgt_seeded_init_(m, 1123)
        m = m.to(dev).train()
        if not sgcls:
            for p in m.object_classifier.parameters():
                p.requires_grad_(False)
        vids = [rank * args.videos + v for v in range(args.videos)]
        batch, (att, spa, con) = _teat_batch(vids, args.frames, (6, 10), dev, sgcls)
        n_pairs = int(batch["pair_idx"].shape[0])
        params = [p for p in m.parameters() if p.requires_grad]
        sync = ddp.GradSync(params[::-1]) if world > 1 else None
        opt = FusedAdamW(params, lr=1e-5, weight_decay=0.1, max_grad_norm=5.0)
        host_ms, host_wait_ms = [], []

        def step():
            m.zero_grad(set_to_none=True)
            pred = m(dict(batch), phase="train")
            loss = (F.cross_entropy(pred["attention_distribution"], att) + F.binary_cross_entropy(pred["spatial_distribution"], spa)
                    + F.binary_cross_entropy(pred["contacting_distribution"], con))
            if sgcls:
                from b200vsgg.objbranch import object_loss
                grp = pred["box_groups"]
                loss = loss + object_loss(pred, 1.0, grp.count, grp.video_of_box64)
            loss = loss + 2500.0 * (pred["structure_temp_loss"].mean() + pred["semantic_temp_loss"].mean())  # TEATGT_train.py:182-185
            loss.backward()
            if sync is not None:
                sync.sync()
            opt.step()
            host_ms.append(m.last_host_graph_ms)
            host_wait_ms.append(getattr(m, "last_host_graph_exposed_ms", m.last_host_graph_ms))
            return loss

        for _ in range(2):
            step()
        steps = max(3, min(args.steps, 5))
        barrier()
        ops.gemm_profile, ops.attn_profile = [], []
        l0 = ops.launch_count
        del host_ms[:], host_wait_ms[:]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        barrier()
        ms = reduce(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None) / steps
        gp, ap = ops.gemm_profile, ops.attn_profile
        ops.gemm_profile = ops.attn_profile = None
        pairs = reduce(float(n_pairs), dist.ReduceOp.SUM if world > 1 else None)
        gf = sum(2.0 * M * N * K for (M, N, K, _, _, _, _) in gp)
        gms = sum(a.elapsed_time(b) for (_, _, _, _, _, a, b) in gp)
        res = {"value": pairs / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "pairs_per_step": int(pairs),
               "tokens_per_gpu": int(m.last_plan.T), "clips_per_gpu": int(m.last_plan.n_clips),
               "gpu_launches_per_step": (ops.launch_count - l0) // steps, "loss": float(loss.item()),
               "host_graph_ms_per_step": float(np.mean(host_ms)),
               "host_graph_wait_ms_per_step": float(np.mean(host_wait_ms)),
               "host_graph_what": "reference-ordered edge-list compaction + LAPACK eigh of the clip Laplacians on the host "
                                  "(parity requires the reference's own eigensolver); SGCls builds it on a worker thread beside "
                                  "the object branch (host_graph_wait_ms = what the main thread still waited for), the "
                                  "single-pass PredCLS step waits for all of it",
               "gemm": {"achieved_tflops": gf / (gms * 1e-3) / 1e12 if gms else 0.0, "frac": gf / (gms * 1e-3) / 1e12 / peak_tf if gms else 0.0,
                        "ms_per_step": gms / steps, "launches": len(gp) // steps}}
        for kind, label in (("fwd", "attention_fwd_tcgen05"), ("bwd", "attention_bwd_tcgen05")):
            fl = sum(f for (k, f, _, _) in ap if k == kind)
            tms = sum(a.elapsed_time(b) for (k, _, a, b) in ap if k == kind)
            res[label] = {"achieved_tflops": fl / (tms * 1e-3) / 1e12 if tms else 0.0,
                          "frac_of_tensor_peak": fl / (tms * 1e-3) / 1e12 / peak_tf if tms else 0.0,
                          "ms_per_step": tms / steps, "launches": sum(1 for a in ap if a[0] == kind) // steps,
                          "bound": "SFU (16 k exp per 128x128 tile = 1024 clk vs ~256 clk of tcgen05.mma at head_dim %d)"
                                   % (768 // targs["encoder_attention_heads"])}
        del m, opt, batch
        torch.cuda.empty_cache()
        return res

    for mode in ("sgcls", "predcls"):
        try:
            out[mode] = run_train(mode)
        except Exception as ex:
            out[mode] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    out["sgcls"]["config"] = "BASELINE configs[2]: TEAT-GT SGCls fwd+bwd, %d videos/GPU x %d frames x 6-10 pairs" % (args.videos, args.frames)
    out["predcls"]["config"] = "TEAT-GT PredCLS fwd+bwd + consistency regulariser, %d videos/GPU x %d frames x 6-10 pairs" % (args.videos, args.frames)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            rate, pairs, threads = teatgt_cpu_rate([0], args.frames, sgcls=True)
            out["sgcls"]["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                            "sample": "1 video (%d pairs) after 1 warm-up pass, oracle/teatgt_oracle.py SGCls "
                                                      "fwd+loss+bwd fp32 train mode" % pairs}
            rate, pairs, threads = teatgt_cpu_rate([0], args.frames, sgcls=False)
            out["predcls"]["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                              "sample": "1 video (%d pairs) after 1 warm-up pass, oracle/teatgt_oracle.py PredCLS "
                                                        "fwd+regulariser+loss+bwd fp32 train mode" % pairs}
        except Exception as ex:
            out["cpu_baseline_error"] = "%s: %s" % (type(ex).__name__, ex)

    if args.longclip or world == 8:
        try:
            m = teatgt.TEAT_GT(mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                               obj_classes=synthetic.ag_object_classes(), tracking=False,
                               args=types.SimpleNamespace(**TEAT_ARGS))
            synthetic.teatgt_seeded_init_(m, 1123)
            m = m.to(dev).eval()
            nv = 8
            batch, _ = _teat_batch([rank * nv + v for v in range(nv)], 256, 32, dev, False)
            n_pairs = int(batch["pair_idx"].shape[0])

            def infer():
                with torch.no_grad():
                    return m(dict(batch), phase="test")

            infer()
            barrier()
            ops.attn_profile = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                pred = infer()
            e1.record()
            barrier()
            ms = reduce(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None) / 2
            ap, ops.attn_profile = ops.attn_profile, None
            fl = sum(f for (_, f, _, _) in ap)
            tms = sum(a.elapsed_time(b) for (_, _, a, b) in ap)
            pairs = reduce(float(n_pairs), dist.ReduceOp.SUM if world > 1 else None)
            # outputs gathered on rank 0 (the only exchange of the config: no data-path collective)
            gathered = None
            if world > 1:
                parts = [torch.empty_like(pred["attention_distribution"]) for _ in range(world)] if rank == 0 else None
                dist.gather(pred["attention_distribution"].contiguous(), parts, dst=0)
                gathered = sum(p.shape[0] for p in parts) if rank == 0 else None
            out["longclip"] = {
                "config": "BASELINE configs[4]: TEAT-GT PredCLS long-clip inference, %d videos x 256 frames x 32 pairs per GPU, "
                          "videos sharded over %d GPU(s), no collective; outputs gathered on rank 0" % (nv, world),
                "value": pairs / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "pairs_per_step": int(pairs),
                "tokens_per_gpu": int(m.last_plan.T), "max_tokens_per_clip": int(m.last_plan.max_T),
                "host_graph_ms_per_step": float(m.last_host_graph_ms), "rows_gathered_on_rank0": gathered,
                "attention_fwd_tcgen05": {"achieved_tflops": fl / (tms * 1e-3) / 1e12 if tms else 0.0,
                                          "frac_of_tensor_peak": fl / (tms * 1e-3) / 1e12 / peak_tf if tms else 0.0,
                                          "ms_per_step": tms / 2, "per_gpu": "rank 0"},
                "reference": "cannot run this config: it keeps 12 x [32,T,T] fp32 attention maps per clip (T = 5337: 43.7 GB)"}
            del m, batch
            torch.cuda.empty_cache()
        except Exception as ex:
            out["longclip"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    return out



def pin_to_gpu_numa_node(local_rank):
    """Run this rank's host thread (and therefore first-touch its pinned staging buffers) on the NUMA node its GPU hangs
    off: at N >= 4 the staging of all ranks otherwise sits on node 0 and the far-socket GPUs copy across the socket
    link (round 1: e2e 67 -> 125 -> 157 ms per step at N = 1 / 4 / 8).  Best effort; returns what was done."""
    try:
        import torch
        props = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            # containers often hide the PCI device's numa_node; NVML still knows which CPUs sit next to the GPU
            try:
                import pynvml
                pynvml.nvmlInit()
                try:
                    h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
                except TypeError:
                    h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
                words = (os.cpu_count() + 63) // 64
                mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
                cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
                allowed = os.sched_getaffinity(0) & cpus
                if allowed and len(allowed) < len(os.sched_getaffinity(0)):
                    os.sched_setaffinity(0, allowed)
                    return {"gpu_pci": bus, "numa_node": node, "pinned": True, "cpus": len(allowed),
                            "how": "nvmlDeviceGetCpuAffinity"}
                return {"gpu_pci": bus, "numa_node": node, "pinned": False,
                        "why": "NVML reports every allowed CPU as local to this GPU (single-node host)"}
            except Exception as ex:
                return {"gpu_pci": bus, "numa_node": node, "pinned": False,
                        "why": "node unknown in sysfs; NVML affinity unavailable (%s)" % type(ex).__name__}
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"gpu_pci": bus, "numa_node": node, "pinned": False, "why": "no allowed CPU on that node"}
        os.sched_setaffinity(0, allowed)
        return {"gpu_pci": bus, "numa_node": node, "pinned": True, "cpus": len(allowed)}
    except Exception as ex:
        return {"pinned": False, "why": "%s: %s" % (type(ex).__name__, ex)}


def workload_config(args, world):
    return {"workload": "TEMPURA PredCLS fwd+bwd, %d synthetic AG videos per GPU (%d frames, 6-10 pairs/frame), "
                        "BASELINE configs[1]%s" % (args.videos, args.frames,
                                                   "" if world == 1 else " x %d GPUs (configs[3] at 8)" % world),
            "videos_per_gpu": args.videos, "frames_per_video": args.frames, "global_videos": args.videos * world,
            "parallelism": "videos sharded over %d GPU(s); gradient all-reduce only" % world,
            "l2": "inputs (3.4 GB/step/GPU) exceed the 126 MB L2",
            "consistency_regulariser": ("off" if getattr(args, "no_consistency", False) else
                                        "on: TEAT-GT regulariser R1-R3 on TEMPURA's graphs, detached as in the reference "
                                        "(lib/teatgt.py:350-351); extension, the reference's TEMPURA never fills these keys"),
            "optimizer_in_step": "fused AdamW + clip_grad_norm_(5) (tools/utils/AdamW.py semantics) after the all-reduce",
            "decoder_row_savings": _decoder_savings_note()}


def _decoder_savings_note():
    """Which of the two result-preserving row savings of the temporal decoder are active (b200vsgg.tempura)."""
    try:
        from b200vsgg import tempura
    except Exception:
        return None
    on = [n for n, f in (("layer-1 q/k/v projections on the N pair rows", tempura.DEC_FIRST_ON_PAIRS),
                         ("last layer after the attention on the N 'latter' rows", tempura.DEC_LATTER_ONLY)) if f]
    return ("; ".join(on) + " (same outputs; model_tflops still counts the dense reference algorithm)") if on else \
        "off (dense M2-row schedule)"


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
TENSOR_KEYS = ("boxes", "labels", "scores", "im_idx", "pair_idx", "human_idx", "features", "union_feat", "union_box",
               "spatial_masks")


def build_batch(video_indices, frames, device):
    """Collated device batch + label tensors + host frame counts (what a loader would hand over)."""
    import numpy as np
    import torch
    from b200vsgg import synthetic, tempura
    entries = [synthetic.make_video_entry(i, frames, (6, 10), device=device, big_on_device=device)
               for i in video_indices]
    gts = [synthetic.build_gt_tensors(e, device) for e in entries]
    batch = tempura.collate_entries(entries)
    for k in ("attention_gt", "spatial_gt", "contacting_gt"):
        batch.pop(k, None)
    batch["gt_tensors"] = tuple(torch.cat([g[i] for g in gts]) for i in range(3))
    batch["frame_counts_host"] = torch.bincount(batch["im_idx"].to(torch.int64)).cpu().numpy().astype(np.int64)
    batch["pair_idx_host"] = batch["pair_idx"].cpu().numpy()      # the loader knows the pairing on the host
    return batch


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from b200vsgg import ddp, ops, synthetic, tempura
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL's CTAs cannot co-reside with the persistent 148-CTA GEMMs (~200 KB of shared memory each): every channel it
        # opens takes an SM away from the GEMM wave that starts next.  16 CTAs move the 0.37 GB of layer buckets well inside
        # the backward they overlap with; measured on one 8 x B200 box, same session (profiles/r02_n8_nccl_ctas.txt):
        # default 47.8 ms/step, 8 CTAs 46.3, 16 CTAs 45.0.
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    numa_info = (pin_to_gpu_numa_node(local_rank) if world > 1 else
                 {"pinned": False, "why": "single rank: the host threads keep every core (CPU baseline runs in this process)"})

    # ---- model (random init of the reference architecture) and this rank's shard of videos
    torch.manual_seed(1123)
    model = tempura.TEMPURA(obj_classes=synthetic.ag_object_classes(), consistency_regulariser=not args.no_consistency,
                            **MODEL_KW)
    synthetic.seeded_init_(model, 1123)
    model = model.to(dev).train()
    for p in model.object_classifier.parameters():  # frozen in PredCLS, TEMPURA_train.py:106-108
        p.requires_grad_(False)
    vids = [rank * args.videos + v for v in range(args.videos)]
    batch = build_batch(vids, args.frames, dev)
    n_pairs = int(batch["pair_idx"].shape[0])
    ddp.broadcast_state(model)            # every rank starts from rank 0's parameters and buffers
    sync = ddp.GradSync(list(model.parameters())[::-1]).attach(model) if world > 1 else None
    # TEMPURA_train.py:111,224-225: AdamW(lr, weight_decay=0.1) after clip_grad_norm_(5) — fused, 2 launches
    from b200vsgg.optim import FusedAdamW
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-5, betas=(0.9, 0.999), eps=1e-8,
                     weight_decay=0.1, max_grad_norm=5.0)

    def run_step(entry):
        model.zero_grad(set_to_none=True)
        pred = model(dict(entry), phase="train")
        losses = tempura.tempura_loss(pred, model.last_plan)
        loss = losses["attention_relation_loss"] + losses["spatial_relation_loss"] + losses["contacting_relation_loss"]
        if "structure_temp_loss" in pred:    # TEMPURA_train.py:215-218 (x2500; detached, so it only shifts the value)
            loss = loss + 2500.0 * (pred["structure_temp_loss"].mean() + pred["semantic_temp_loss"].mean())
        loss.backward()
        if sync is not None:
            sync.sync()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ============================== device-resident timing ==============================
    if args.profile:
        args.no_e2e = args.no_cpu_baseline = True
    for _ in range(1 if args.profile else max(args.warmup, 3)):
        run_step(batch)
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    ops.gemm_profile = []
    launches0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = run_step(batch)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count - launches0
    gemm_prof, ops.gemm_profile = ops.gemm_profile, None
    clock_info = clocks.stop()
    comm = None
    if sync is not None:
        # what the post-backward exchange costs ON the compute stream: waits for the layer-bucket all-reduces issued
        # during backward + the bucketed rest (copy in, all-reduce, copy out); everything else overlapped
        ev = sync.comm_events[-args.steps:]
        exposed = sum(a.elapsed_time(b) for a, b in ev) / max(1, len(ev))
        comm = {"comm_exposed_ms_per_step": max_over_ranks(exposed),
                "layer_buckets": len(sync._layer_flat), "layer_bucket_mb": sum(t.numel() * 4 for t in sync._layer_flat) / 1e6,
                "note": "CUDA events around GradSync.sync() on the compute stream, max over ranks"}
        sync.comm_events = []
        ddp.sync_buffers(model)           # BatchNorm running statistics: averaged over the ranks (outside the timed region)
    total_pairs = sum_over_ranks(float(n_pairs))
    value = total_pairs * args.steps / (ms_total * 1e-3)
    loss_val = float(loss.item())

    # ---- roofline of the dominant kernel from the per-launch CUDA events
    flops = sum(2.0 * M * N * K for (M, N, K, _, _, _, _) in gemm_prof)
    gemm_ms = sum(a.elapsed_time(b) for (_, _, _, _, _, a, b) in gemm_prof)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    if peak is None:
        peak, peak_src = 1400.0, "fallback of B200_PROFILING.md (sustained ~1.4 PFLOP/s)"
    achieved = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "vsgg::gemm_bf16_kernel (tcgen05/TMA, all instantiations)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "launches_timed": len(gemm_prof),
                "gemm_ms_per_step": gemm_ms / args.steps, "gemm_share_of_step": gemm_ms / (e0.elapsed_time(e1)),
                "gemm_flops_per_step": flops / args.steps}
    # DRAM traffic of the kernel from the committed `ncu --set full` capture (never measured under this run)
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_gemm_traffic.json")) as f:
            tr = json.load(f)
        rep = max(tr["launches"], key=lambda l: l["shape"][0] * l["shape"][1] * l["shape"][2])
        roofline["traffic"] = rep["traffic_bytes"] / 1e9
        roofline["traffic_detail"] = {
            "unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "shape_MNK": rep["shape"],
            "algorithmic_GB": rep["algorithmic_bytes"] / 1e9, "traffic_over_algorithmic": rep["traffic_over_algorithmic"],
            "all_captured": [[l["shape"], round(l["traffic_over_algorithmic"], 3)] for l in tr["launches"]],
            "source": "profiles/r01_ncu_gemm_traffic.json (" + tr["source"] + ")"}
    except Exception:
        pass
    if args.gemm_log and rank == 0:
        agg = {}
        for (M, N, K, a_mn, b_mn, a, b) in gemm_prof:
            r = agg.setdefault((M, N, K, a_mn, b_mn), [0, 0.0])
            r[0] += 1
            r[1] += a.elapsed_time(b)
        with open(args.gemm_log, "w") as f:
            for (M, N, K, a_mn, b_mn), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(json.dumps({"M": M, "N": N, "K": K, "a_mn": a_mn, "b_mn": b_mn, "launches": cnt,
                                    "ms_total": ms, "ms_avg": ms / cnt,
                                    "tflops": 2.0 * M * N * K * cnt / (ms * 1e-3) / 1e12}) + "\n")

    # ============================== end-to-end (host inputs) ==============================
    # Two hand-off contracts, same step, same videos:
    #   reference_fp32_nchw   what tools/utils/object_detector.py:372-396 produces today: union_feat fp32 [N,1024,7,7],
    #                         spatial_masks fp32 [N,2,27,27] (3.55 GB per 64-video step: PCIe-bound)
    #   producer_bf16_nhwc    SURVEY.md §8 (f).4: the detector emits what the path consumes — union_feat bf16 channels-
    #                         last rows, bf16 masks (tempura.to_producer_contract): half the bytes, no layout kernel,
    #                         same bf16 operands in every GEMM (tests/test_tempura_gpu.py)
    # `e2e` (the contract's headline key) is the reference contract; `e2e_producer_contract` sits beside it.
    e2e = e2e_fast = None
    if not args.no_e2e:
        fast_dev = tempura.to_producer_contract({"union_feat": batch["union_feat"], "spatial_masks": batch["spatial_masks"]})
        fast_host = {k: fast_dev[k].cpu().pin_memory() for k in ("union_feat", "spatial_masks")}
        del fast_dev
        host = {k: batch[k].cpu().pin_memory() for k in TENSOR_KEYS if k in batch}
        gt_host = tuple(t.cpu().pin_memory() for t in batch["gt_tensors"])
        meta = {k: batch[k] for k in ("video_frames", "frame_counts_host", "pair_idx_host", "video_size")}
        gt_like = batch["gt_tensors"]
        small_like = {k: batch[k] for k in host if k not in ("union_feat", "spatial_masks")}
        del batch["union_feat"], batch["spatial_masks"]
        torch.cuda.empty_cache()
        copy_stream = torch.cuda.Stream(device=dev)

        def run_e2e(host_set):
            h2d = sum(t.numel() * t.element_size() for t in host_set.values()) + sum(t.numel() * t.element_size() for t in gt_host)
            bufs = [({k: torch.empty(t.shape, dtype=t.dtype, device=dev) for k, t in host_set.items()},
                     tuple(torch.empty_like(t) for t in gt_like)) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]

            def issue_copy(slot):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[slot])
                    dst, gdst = bufs[slot]
                    for k, t in host_set.items():
                        # one cudaMemcpyAsync per 64 MB slice (index vectors no longer use the copy engine: the model
                        # uploads them with a kernel that reads pinned memory, so nothing small queues behind these)
                        n, chunk = t.numel(), (64 << 20) // t.element_size()
                        if n <= chunk:
                            dst[k].copy_(t, non_blocking=True)
                        else:
                            df, sf = dst[k].view(-1), t.view(-1)
                            for i in range(0, n, chunk):
                                df[i:i + chunk].copy_(sf[i:i + chunk], non_blocking=True)
                    for d, s_ in zip(gdst, gt_host):
                        d.copy_(s_, non_blocking=True)
                    ready[slot].record(copy_stream)

            def loop(n):
                for s_ in range(2):
                    consumed[s_].record()
                issue_copy(0)
                last = None
                for i in range(n):
                    slot = i & 1
                    if i + 1 < n:
                        issue_copy(slot ^ 1)
                    torch.cuda.current_stream().wait_event(ready[slot])
                    dst, gdst = bufs[slot]
                    entry = dict(dst)
                    entry.update(meta)
                    entry["gt_tensors"] = gdst
                    loss_ = run_step(entry)
                    consumed[slot].record()
                    last = float(loss_.item())  # device -> host read of the step's result
                return last

            loop(max(args.warmup, 3))
            barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            loop(args.steps)
            t1.record()
            barrier()
            ms = max_over_ranks(t0.elapsed_time(t1))
            del bufs
            torch.cuda.empty_cache()
            return {"value": total_pairs * args.steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms / args.steps,
                    "h2d_gb_per_s_per_gpu": h2d / (ms / args.steps * 1e-3) / 1e9}

        e2e = run_e2e(host)
        e2e["contract"] = "reference_fp32_nchw (tools/utils/object_detector.py:372-396 as is)"
        e2e["how"] = "model.forward(entry)+loss+backward+clip+AdamW on pinned-host inputs, H2D double-buffered on a copy stream"
        host_fast = dict(host)
        host_fast.update(fast_host)
        del host["union_feat"], host["spatial_masks"]
        e2e_fast = run_e2e(host_fast)
        e2e_fast["contract"] = "producer_bf16_nhwc (SURVEY 8(f).4: union_feat bf16 [N,7,7,1024], masks bf16; same bf16 operands in every GEMM)"
        e2e_fast["numa"] = numa_info

    # ============================== CPU baseline (rank 0, N = 1) ==============================
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_vids = vids[:args.cpu_sample_videos]
        rate, sec, pairs, cores, threads = cpu_reference_rate(sample_vids, args.frames, steps=2, warmup=1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "host_cores": cores,
               "sample": "%d of the step's %d videos (%d pairs), 2 timed passes after 1 warm-up, oracle/tempura_oracle.py "
                         "fwd+loss+bwd fp32 train mode, one video per forward" % (len(sample_vids), args.videos, pairs)}

    # ============================== library baseline (rank 0, N = 1) ==============================
    lib = None
    if rank == 0 and world == 1 and not args.no_library_baseline and not args.profile:
        try:
            lib = library_baseline_rates(vids[:args.lib_sample_videos], args.frames, dev)
            for k in ("eager_fp32_tf32", "eager_bf16_autocast"):
                lib[k]["b200_value_over_this"] = value / lib[k]["value"]
                if e2e is not None:
                    lib[k]["b200_e2e_over_this"] = e2e["value"] / lib[k]["value"]
        except Exception as ex:  # the baseline is a report, never a reason to lose the headline line
            lib = {"error": "%s: %s" % (type(ex).__name__, ex)}

    # ============================== TEAT-GT section (all ranks) ==============================
    teat = None
    if not args.no_teatgt and not args.profile:
        del batch
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        try:
            teat = teatgt_section(args, dev, rank, world, peaks)
        except Exception as ex:
            teat = {"error": "%s: %s" % (type(ex).__name__, ex)}

    if rank == 0:
        cfg = workload_config(args, world)
        cfg["pairs_per_step"] = int(total_pairs)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
                "clocks": clock_info, "e2e": e2e, "e2e_producer_contract": e2e_fast, "comm": comm,
                "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu, "library_baseline": lib, "teatgt": teat, "loss": loss_val,
                "model_tflops": 1.138e9 * value / 1e12}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
