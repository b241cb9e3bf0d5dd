"""Per bench-class DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) from an `ncu --set full`
capture of bench.py, keyed like bench.py's kernel classes, for `roofline.traffic`:
    python tools/ncu_traffic.py profiles/ncu_rNN_traffic.json gpurun_out/prof_fwd.ncu-rep [gpurun_out/prof_bwd.ncu-rep ...]
Forward kernels precede the first wgrad (gemm_tc_kernel<4,0>) in a capture; after it <0,1>/<3,1> are dgrads.  A capture
whose file name contains "bwd" is all backward."""
import collections
import csv
import io
import json
import subprocess
import sys


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main(out, reps):
    acc = collections.defaultdict(list)
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        rd, wr, nm = ix["dram__bytes_read.sum"], ix["dram__bytes_write.sum"], ix["Kernel Name"]
        backward = "bwd" in rep
        for r in data:
            name = r[nm]
            if "gemm_tc_kernel<4, 0" in name:
                backward = True
            if "gemm_tc_kernel<4, 0" in name: cls = "gemm_wgrad"
            elif "mlp_tc_kernel<0" in name: cls = "gemm_mlp_fwd"
            elif "mlp_tc_kernel<1" in name: cls = "gemm_mlp_bwd"
            elif "gemm_tc_kernel<3, 1" in name: cls = "gemm_dgrad"
            elif "gemm_tc_kernel<0, 1" in name: cls = "gemm_dgrad" if backward else "gemm_qkv"
            elif "gemm_tc_kernel<2, 1" in name: cls = "gemm_fc1"
            elif "gemm_tc_kernel<1, 0" in name: cls = "gemm_proj"           # fc2 lives in the chained MLP kernel since round 2
            elif "attn_fwd" in name: cls = "attn_fwd"
            elif "attn_bwd" in name: cls = "attn_bwd"
            elif "ln_bwd" in name: cls = "ln_bwd"
            else: continue
            acc[cls].append(to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]))
    res = {k: {"launches_captured": len(v), "dram_bytes_per_launch": sum(v) / len(v)} for k, v in acc.items()}
    res["_source"] = (f"ncu --set full --clock-control none, {' + '.join(reps)} (dram__bytes_read.sum + dram__bytes_write.sum, mean "
                      "over captured launches; cold-cache, serialised; write-back of a kernel's output may be charged to a later launch)")
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
