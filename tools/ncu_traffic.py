"""profiles/roofline_traffic.json from an `ncu --set full` capture of one eager train step: mean DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) per launch of each conv kernel, keyed by the C-ABI entry point bench.py
reports as `roofline.kernel`.  Usage: python tools/ncu_traffic.py capture.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys

# keyed by KERNEL name: bench.py aggregates every entry point that launches the same kernel into one roofline population
# (conv3d_wgrad_tc_kernel = the k3 / k(1,3,3) wgrads + head wgrad + the stem's k(3,1,1) wgrad: all its launches of one step)
WG = "conv3d_wgrad_hs_kernel+conv3d_wgrad_tc_kernel"      # (+ conv3d_wgrad_rs_kernel, the head)
_WG_DOC = WG      # bench.py WGRAD_KERNELS: the weight-gradient kernels, one population
ENTRY = {"conv3d_wgrad_tc_kernel": WG, "conv3d_wgrad_hs_kernel": WG, "conv3d_wgrad_rs_kernel": WG, "conv3d_tc_dfold_kernel": "conv3d_tc_dfold_kernel",
         "conv3d_tc_kernel": "conv3d_tc_kernel"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else "profiles/roofline_traffic.json"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi, ti = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                             "gpu__time_duration.sum"))
    agg = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "").split("<")[0]
        if name not in ENTRY:
            continue
        b = float(r[ri].replace(",", "")) * UNIT[units[ri]] + float(r[wi].replace(",", "")) * UNIT[units[wi]]
        a = agg.setdefault(ENTRY[name], {"kernel": name, "launches": 0, "bytes": 0.0, "us": 0.0})
        a["launches"] += 1
        a["bytes"] += b
        a["us"] += float(r[ti].replace(",", ""))
    res = {k: {"kernel": v["kernel"], "launches_captured": v["launches"], "dram_bytes_per_launch": v["bytes"] / v["launches"],
               "mean_us_under_ncu": v["us"] / v["launches"], "source": rep} for k, v in agg.items()}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
