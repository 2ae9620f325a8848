"""Where does the device wait?  Runs the benchmarked TEMPURA step (bench.py's run_step, BASELINE configs[1]) under the
CUPTI kernel tracer of torch.profiler and prints, for one steady-state step: device busy time, idle time, and the
largest idle gaps with the kernels either side of each gap (the kernel AFTER a gap names the host work that was late).
Diagnostic only: nothing measured under a profiler is ever reported as a bench value.

    python tools/step_timeline.py [--videos 64] [--frames 32] [--top 25] [--out gpurun_out/timeline.txt]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=64)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-consistency", action="store_true")
    args = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
    import bench
    from b200vsgg import synthetic, tempura
    from b200vsgg.optim import FusedAdamW
    dev = torch.device("cuda", 0)
    torch.manual_seed(1123)
    model = tempura.TEMPURA(obj_classes=synthetic.ag_object_classes(), consistency_regulariser=not args.no_consistency,
                            **bench.MODEL_KW)
    synthetic.seeded_init_(model, 1123)
    model = model.to(dev).train()
    for p in model.object_classifier.parameters():
        p.requires_grad_(False)
    batch = bench.build_batch(list(range(args.videos)), args.frames, dev)
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-5, betas=(0.9, 0.999), eps=1e-8,
                     weight_decay=0.1, max_grad_norm=5.0)

    def run_step():
        model.zero_grad(set_to_none=True)
        pred = model(dict(batch), phase="train")
        losses = tempura.tempura_loss(pred, model.last_plan)
        loss = losses["attention_relation_loss"] + losses["spatial_relation_loss"] + losses["contacting_relation_loss"]
        if "structure_temp_loss" in pred:
            loss = loss + 2500.0 * (pred["structure_temp_loss"].mean() + pred["semantic_temp_loss"].mean())
        loss.backward()
        opt.step()
        return loss

    for _ in range(4):
        run_step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=False) as prof:
        for _ in range(args.steps):
            run_step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps at the optimiser kernel
    ends = [i for i, e in enumerate(evs) if "adamw_clip" in e.name]
    lines = []
    if len(ends) >= 2:
        a, b = ends[-2] + 1, ends[-1] + 1
    else:
        a, b = 0, len(evs)
    step = evs[a:b]
    t0, t1 = step[0].time_range.start, step[-1].time_range.end
    busy, cur_end, gaps = 0.0, t0, []
    for i, e in enumerate(step):
        s, en = e.time_range.start, e.time_range.end
        if s > cur_end:
            gaps.append((s - cur_end, step[i - 1].name if i else "-", e.name, cur_end, s))
            busy += en - s
        else:
            busy += max(0.0, en - cur_end)
        cur_end = max(cur_end, en)
    span = t1 - t0
    lines.append("one steady-state step: %d device activities, span %.3f ms, busy %.3f ms, idle %.3f ms (%d gaps)"
                 % (len(step), span / 1e3, busy / 1e3, (span - busy) / 1e3, len(gaps)))
    small = sum(g[0] for g in gaps if g[0] < 5.0)
    lines.append("idle in gaps < 5 us (launch-to-launch latency): %.3f ms over %d gaps; >= 5 us: %.3f ms over %d gaps"
                 % (small / 1e3, sum(1 for g in gaps if g[0] < 5.0), (span - busy - small) / 1e3,
                    sum(1 for g in gaps if g[0] >= 5.0)))
    lines.append("largest gaps (us)   kernel before  ->  kernel after")
    cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU]
    for g, p, n, ga, gb in sorted(gaps, key=lambda x: -x[0])[:args.top]:
        lines.append("  %8.1f   %-60s -> %s" % (g, p[:60], n[:80]))
        if g >= 30.0:
            # host activity during the gap: the longest CPU-side events that overlap it (what the host was busy with)
            ov = {}
            for c in cpu:
                a, b = c.time_range.start, c.time_range.end
                o = min(b, gb) - max(a, ga)
                if o > 0:
                    r = ov.setdefault(c.name[:60], [0.0, 0])
                    r[0] += o
                    r[1] += 1
            for name, (o, cnt) in sorted(ov.items(), key=lambda kv: -kv[1][0])[:8]:
                lines.append("               host: %-60s x%-4d overlap %.0f us" % (name, cnt, o))
    # aggregate gap time by the kernel that follows
    agg = {}
    for g, p, n, _, _ in gaps:
        k = n[:70]
        r = agg.setdefault(k, [0.0, 0])
        r[0] += g
        r[1] += 1
    lines.append("gap time by following kernel (us total, count)")
    for k, (g, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:args.top]:
        lines.append("  %8.1f  x%-4d %s" % (g, c, k))
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
