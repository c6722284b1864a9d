"""Where the replayed step idles: kernel timeline of CUDA-graph replays of the bench step (CUPTI through torch.profiler).

    python tools/gap_profile.py [out.txt]        # XFM_PROFILE_GRAPH=0 profiles the launch-by-launch sequence instead

Per replay: span (first kernel start -> last kernel end), busy time (union of the kernel intervals over all streams), the
idle remainder, the histogram of the gaps between consecutive kernels and the kernels that most often FOLLOW a gap.  The
idle total is the upper bound of what programmatic dependent launch / fewer launches could still recover.
"""
import json
import os
import sys
import tempfile
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from xfm_b200.accelerator import B200DDPAccelerator, FlatAdamW  # noqa: E402
from xfm_b200.graph import GraphedStep  # noqa: E402
from xfm_b200.model_pretrain import XFM  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gap_profile.txt"
    use_graph = os.environ.get("XFM_PROFILE_GRAPH", "1") != "0"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    model = XFM(bench.base_config(), init=bench.gpu_init(dev, 0), device=dev).train()
    opt = FlatAdamW(model, lr=1e-4, weight_decay=0.01, lr_mult=2.0)
    acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0))
    b = {k: v.to(dev) for k, v in bench.make_host_batch(96, 40, 15, model.cfg["vocab_size"], 224, 100).items()}

    def loss_fn(m, b):
        out = m(b["image"], b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
                masked_ids=b["masked_ids"], ret_mim_loss=True, data_source="image")
        return out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + out["loss_mim"]

    if use_graph:
        gs = GraphedStep(model, opt, acc, loss_fn, b, warmup=3, uses_mim_masks=True)
        step = lambda: gs(b)  # noqa: E731
    else:
        def step():
            loss = loss_fn(model, b)
            acc.backward_step(loss, opt)
            acc.optimizer_step(opt, model)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n_rep = 3
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for _ in range(n_rep):
            step()
            torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    # split into replays at the largest n_rep - 1 gaps (the host synchronises between steps)
    gaps = sorted(((ev[i + 1]["ts"] - (ev[i]["ts"] + ev[i]["dur"]), i) for i in range(len(ev) - 1)), reverse=True)[:n_rep - 1]
    cuts = sorted(i + 1 for _, i in gaps)
    parts = [ev[a:b_] for a, b_ in zip([0] + cuts, cuts + [len(ev)])]
    lines = [f"# {'CUDA-graph replay' if use_graph else 'launch-by-launch sequence'} of the bench step, {n_rep} steps, "
             f"{len(ev)} device activities (CUPTI via torch.profiler; times in us)"]
    for k, p in enumerate(parts):
        span = p[-1]["ts"] + p[-1]["dur"] - p[0]["ts"]
        busy, end, hist, after = 0.0, p[0]["ts"], defaultdict(lambda: [0, 0.0]), defaultdict(lambda: [0, 0.0])
        for i, e in enumerate(p):
            s, t = e["ts"], e["ts"] + e["dur"]
            if s > end:
                g = s - end
                bucket = "<1" if g < 1 else "1-2" if g < 2 else "2-4" if g < 4 else "4-8" if g < 8 else ">=8"
                hist[bucket][0] += 1
                hist[bucket][1] += g
                name = e["name"].split("<")[0].split("(")[0][-48:]
                after[name][0] += 1
                after[name][1] += g
                busy += t - s
                end = t
            elif t > end:
                busy += t - end
                end = t
        lines.append(f"step {k}: {len(p)} activities, span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms "
                     f"({100 * (span - busy) / span:.1f} %), sum of durations {sum(e['dur'] for e in p) / 1e3:.2f} ms")
        lines.append("  gap histogram (us): " + ", ".join(f"{b_}: {hist[b_][0]} gaps / {hist[b_][1] / 1e3:.2f} ms" for b_ in ("<1", "1-2", "2-4", "4-8", ">=8")))
        if k == len(parts) - 1:
            lines.append("  idle time by the kernel that follows the gap (top 15):")
            for name, (cnt, tot) in sorted(after.items(), key=lambda kv: -kv[1][1])[:15]:
                lines.append(f"    {tot:9.1f} us  {cnt:5d} gaps  {tot / cnt:6.2f} us/gap  {name}")
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
