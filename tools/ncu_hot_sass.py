"""Top SASS instructions by warp-stall samples from an .ncu-rep captured with --import-source on (run here, no GPU):
    python tools/ncu_hot_sass.py gpurun_out/prof.ncu-rep [top_n] [kernel_index]"""
import csv
import io
import subprocess
import sys


def main(path, top=40, which=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name",')
    blk = blocks[1 + which]
    lines = blk.split("\n")
    print("# kernel:", lines[0][:150])
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[1:] if len(r) == len(hdr)]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(f"# {len(data)} SASS instructions, {tot} samples; total executed warp-instr {sum(int(r[ix['Instructions Executed']] or 0) for r in data)}")
    agg = {}
    for r in data:
        for h in stall_cols:
            agg[h] = agg.get(h, 0) + int(r[ix[h]] or 0)
    print("# stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
    for i in sorted(order):
        r = data[i]
        st = {h[6:]: int(r[ix[h]] or 0) for h in stall_cols if int(r[ix[h]] or 0) > 0}
        top2 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"{i:5d} {int(r[ix['# Samples']]):7d} {100.0 * int(r[ix['# Samples']]) / tot:5.1f}%  exec {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:70]:70s} {top2}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
