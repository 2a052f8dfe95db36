"""Per-source-line summary of one launch of an `ncu --set full --import-source on` capture:
    python profiles/tools/ncu_lines.py <capture.ncu-rep> <launch index> [top N]
Lists the source lines with the most warp-stall samples (share of the kernel), their executed warp instructions, shared-memory
wavefronts, global L1 tag requests and dominant stall reasons."""
import csv
import subprocess
import sys


def main(path, skip, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--launch-skip", str(skip), "--launch-count", "1", "--page", "source", "--csv",
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    fname, hdr, lines = None, None, []
    kernel = ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            kernel = r[1].split("(")[0].split("::")[-1] + r[1].split("k_mf_")[1].split("(msfec")[0] if "k_mf_" in r[1] else r[1][:60]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr and len(r) == len(hdr) and r[2] == "-":
            lines.append((fname, dict(zip(hdr[4:], r[4:])), r[0], r[1]))
    def num(d, k):
        try:
            return float(d[k])
        except Exception:
            return 0.0
    tot = sum(num(d, "# Samples") for _, d, _, _ in lines) or 1.0
    tot_i = sum(num(d, "Instructions Executed") for _, d, _, _ in lines) or 1.0
    tot_w = sum(num(d, "L1 Wavefronts Shared") for _, d, _, _ in lines) or 1.0
    tot_g = sum(num(d, "L1 Tag Requests Global") for _, d, _, _ in lines) or 1.0
    print(f"kernel {kernel}: {tot:.0f} samples, {tot_i:.3g} warp instr, {tot_w:.3g} smem wavefronts, {tot_g:.3g} global tag requests")
    stall_keys = [k for k in lines[0][1] if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(num(d, k) for _, d, _, _ in lines) for k in stall_keys}
    print("stalls:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    lines.sort(key=lambda t: -num(t[1], "# Samples"))
    print("| file:line | samples % | instr % | smem wavefronts % | global tags % | top stalls | source |")
    print("|---|---|---|---|---|---|---|")
    for f, d, ln, src in lines[:top]:
        st = sorted(((num(d, k), k[6:]) for k in stall_keys), reverse=True)[:2]
        print(f"| {f}:{ln} | {100 * num(d, '# Samples') / tot:.1f} | {100 * num(d, 'Instructions Executed') / tot_i:.1f} | "
              f"{100 * num(d, 'L1 Wavefronts Shared') / tot_w:.1f} | {100 * num(d, 'L1 Tag Requests Global') / tot_g:.1f} | "
              + ", ".join(f"{k} {100 * v / max(num(d, '# Samples'), 1):.0f}%" for v, k in st) + f" | `{src.strip()[:110]}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 40)
