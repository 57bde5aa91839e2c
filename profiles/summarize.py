"""Turns the raw outputs of a gpurun measurement call into the tracked summaries of this directory.

    python profiles/summarize.py gpurun_out r1

reads  <dir>/bench.json (bench.py stdout), <dir>/step_profile.json (bench.py --profile-out), <dir>/launches.csv (ncu
--metrics gpu__time_duration.sum --csv log) and writes  profiles/<tag>_bench.json, <tag>_step_profile.json/.md and
<tag>_launches_summary.md."""
import csv
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def short(name: str) -> str:
    name = name.replace("void ", "")
    m = re.match(r"(?:vsb::)?(conv_gemm_kernel<[^>]*>|conv3_direct_kernel<[^>]*>|[A-Za-z0-9_:]+(?:<[0-9, a-z]*>)?)", name)
    return (m.group(1) if m else name)[:80]


def main(src: str, tag: str):
    bench = json.loads(open(os.path.join(src, "bench.json")).read().strip().splitlines()[-1])
    json.dump(bench, open(os.path.join(HERE, f"{tag}_bench.json"), "w"))
    prof = json.load(open(os.path.join(src, "step_profile.json")))
    json.dump(prof, open(os.path.join(HERE, f"{tag}_step_profile.json"), "w"), indent=1)
    rows = prof["table"]
    total = sum(r["ms_per_step"] for r in rows)
    with open(os.path.join(HERE, f"{tag}_step_profile.md"), "w") as f:
        f.write(f"# {tag} — per-kernel step profile (CUDA events around every plan step inside `bench.py`, batch {prof['batch']}, {prof['card']})\n\n")
        f.write(f"Step under events: {total:.2f} ms (timed step without events: {bench['ms_per_step']:.2f} ms = {bench['value']:.0f} frames/s). "
                f"Source: `profiles/{tag}_step_profile.json` (bench.py --profile-out).\n\n")
        f.write("| plan step (shape tag = C_in-C_out@map) | ms/step | launches | avg µs | TFLOP/s | share |\n|---|---|---|---|---|---|\n")
        for r in sorted(rows, key=lambda r: -r["ms_per_step"]):
            tf = f"{r['tflops']:.0f}" if r.get("tflops") else ""
            f.write(f"| `{r['name']}` | {r['ms_per_step']:.3f} | {r['launches_per_step']} | {r['avg_us']:.1f} | {tf} | {100 * r['ms_per_step'] / total:.1f}% |\n")
    lc = os.path.join(src, "launches.csv")
    if os.path.exists(lc):
        lines = [l for l in open(lc) if l.startswith('"')]
        recs = list(csv.DictReader(lines))
        ours = [r for r in recs if r["Metric Name"] == "gpu__time_duration.sum"]
        # the timed steps are the tail of the run: keep the launches of the last 2 of 5 steps
        per_step = bench.get("gpu_launches_per_step") or None
        steps_total = 5
        tail = ours[-(len(ours) * 2 // steps_total):] if not per_step else ours[-2 * per_step:]
        ev = prof["table"]
        ev_tot = sum(r["ms_per_step"] for r in ev)
        ev_gemm = sum(r["ms_per_step"] for r in ev if r.get("tflops")) / ev_tot if ev_tot else 0
        agg = {}
        for r in tail:
            k = short(r["Kernel Name"])
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += float(r["Metric Value"]) / 1e6
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(HERE, f"{tag}_launches_summary.md"), "w") as f:
            f.write(f"# {tag} - ncu launch list of `bench.py --steps 2 --warmup 3` (gpu__time_duration.sum, --clock-control none)\n\n")
            f.write(f"{len(tail)} launches = the last 2/5 of the {len(ours)} captured (cold-cache, serialised: compare SHARES). Total {tot:.2f} ms.\n\n")
            f.write(f"(event profile of the same command: the tensor-core conv / GEMM steps are {100 * ev_gemm:.1f} % of the step)\n\n")
            f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
