"""Per-plan-step roofline table from a bench.py --profile-out file: algorithmic FLOPs and minimum HBM bytes of every step
(from its shape tag, videoseal_b200/csrc/model.cuh step names) -> ideal time = max(flops / tensor peak, bytes / HBM peak) ->
achieved fraction.  python profiles/layer_roofline.py profiles/r1_step_profile.json MEASURED_PEAKS.json > profiles/r1_layer_roofline.md"""
import json
import sys


def model(name, B):
    """(flops, bytes) of ONE launch; activations fp16 NHWC unless noted; weights ignored when << activations"""
    kind, _, rest = name.rpartition(".")
    dims, _, hs = rest.partition("@")
    try:
        h = int(hs)
    except ValueError:
        return None
    M = B * h * h
    f2, f4 = 2, 4
    ints = [int(x) for x in dims.split("-")] if dims.replace("-", "").isdigit() else []
    if kind in ("unet.conv3x3", "unet.conv3x3d"):
        ci, co = ints
        return 2.0 * M * co * 9 * ci, M * (ci * f2 + co * f2 + co * f2) + 9 * ci * co * f2      # in, out, residual or none (upper bound)
    if kind in ("unet.conv3x3+outc", "unet.conv3x3d+outc"):
        ci, co = ints
        return 2.0 * M * co * 9 * ci, M * (ci * f2 + co * f2 + 4)
    if kind in ("unet.conv1x1", "unet.conv1x1d"):
        ci, co = ints
        return 2.0 * M * co * ci, M * (ci + co) * f2 + ci * co * f2
    if kind in ("unet.down3x3s2", "unet.down3x3s2d"):
        ci, co = ints
        return 2.0 * M * co * 9 * ci, M * (4 * ci + co) * f2
    if kind == "unet.uptap1x1":
        ci, co = ints
        return 2.0 * M * co * ci, M * (ci + co) * f2 + ci * co * f2
    if kind == "unet.upgather":
        co = ints[0]
        return 0.0, (M // 4) * 9 * co * f2 + M * co * f2          # low-res tap tensor read once, output written
    if kind == "unet.upphase":
        ci, co = ints
        return 2.0 * (M // 4) * 4 * co * 9 * ci, (M // 4) * ci * f2 + M * co * f2
    if kind == "cnx.dwconv7_ln":
        c = ints[0]
        return 2.0 * M * c * 49, M * c * (f4 + f2)
    if kind == "cnx.pwconv1":
        c = ints[0]
        return 2.0 * M * c * 4 * c, M * (c + 4 * c) * f2
    if kind == "cnx.pwconv2":
        c = ints[0]
        return 2.0 * M * c * 4 * c, M * (4 * c * f2 + 2 * c * f4)
    if kind == "cnx.grn_apply":
        c = ints[0]
        return 0.0, M * 4 * c * f2 * 2
    if kind == "cnx.down2x2s2":
        ci, co = ints
        return 2.0 * M * co * 4 * ci, M * (4 * ci * f2 + co * f4)
    if kind == "cnx.head3x3":
        c = ints[0]
        return 2.0 * M * c * 9 * c, M * c * (f2 + f4) + 9 * c * c * f2
    if kind == "cnx.stem_gemm":
        c = ints[0]
        return 2.0 * M * c * 48, M * (64 * f2 + c * f4)
    return None


def main(prof_path, peaks_path):
    prof = json.load(open(prof_path))
    peaks = json.load(open(peaks_path))
    tf, bw = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) * 1e12, peaks["hbm_gbs"] * 1e9
    B = prof["batch"]
    rows, tot_meas, tot_ideal, tot_unmodelled = [], 0.0, 0.0, 0.0
    for r in prof["table"]:
        m = model(r["name"], B)
        meas = r["ms_per_step"] * 1e-3
        tot_meas += meas
        if not m:
            tot_unmodelled += meas
            continue
        fl, by = m
        ideal1 = max(fl / tf, by / bw)
        ideal = ideal1 * r["launches_per_step"]
        tot_ideal += ideal
        rows.append((r["name"], r["launches_per_step"], r["avg_us"], ideal1 * 1e6, "tensor" if fl / tf > by / bw else "hbm",
                     ideal / meas, meas * 1e3))
    print(f"# Per-step roofline model vs measurement ({prof['card']}, batch {B}; peaks: {tf/1e12:.0f} TF/s sustained, {bw/1e9:.0f} GB/s)\n")
    print("Ideal time of a launch = max(algorithmic FLOPs / tensor peak, minimum HBM bytes / HBM peak); weights and L2 reuse ignored.\n")
    print("| plan step | launches | measured µs | ideal µs | bound | achieved fraction | ms/step |\n|---|---|---|---|---|---|---|")
    for n, l, us, ius, bd, fr, ms in sorted(rows, key=lambda t: -t[6]):
        print(f"| `{n}` | {l} | {us:.1f} | {ius:.1f} | {bd} | {fr:.2f} | {ms:.3f} |")
    print(f"\nModelled steps: measured {1e3 * (tot_meas - tot_unmodelled):.2f} ms vs ideal {1e3 * tot_ideal:.2f} ms per step "
          f"(overall fraction {tot_ideal / (tot_meas - tot_unmodelled):.2f}); un-modelled pointwise steps: {1e3 * tot_unmodelled:.2f} ms.")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
