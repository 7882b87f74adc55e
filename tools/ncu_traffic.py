#!/usr/bin/env python
"""profiles/ncu_traffic.json from one `ncu --set full` capture of a single scene pair:
dram__bytes_read.sum + dram__bytes_write.sum of the kernels of each bench.py stage (first launch
of each kernel; both Laplacian launches).  bench.py reports the dominant stage's figure as
roofline.traffic."""
import csv, io, json, subprocess, sys

STAGES = [("minmax_mask", ["k_minmax_mask"]), ("laplacian_mon", ["k_laplacian4"]), ("laplacian_ref", ["k_laplacian4"]),
          ("corner_response", ["k_eig_approx", "k_exact_max"]),
          ("select", ["k_cand_hist_fast", "k_cutoff", "k_select", "k_exact_cands"]),
          ("nms", ["k_nms"]), ("pyramids", ["k_pyr_down4"]), ("lk_roundtrip", ["k_lk_roundtrip"]),
          ("zncc", ["k_zncc"])]


def main(rep, source_note, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def nbytes(r, k):
        return float(r[ix[k]].replace(",", "")) * scale[units[ix[k]].strip()]
    launches = [(r[ix["Kernel Name"]], nbytes(r, "dram__bytes_read.sum") + nbytes(r, "dram__bytes_write.sum"),
                 r[ix["gpu__time_duration.sum"]] + units[ix["gpu__time_duration.sum"]]) for r in data]
    used, stages = set(), {}
    for name, kernels in STAGES:
        tot, names = 0.0, []
        for k in kernels:
            for i, (kn, b, d) in enumerate(launches):
                if i not in used and k in kn:
                    used.add(i); tot += b; names.append(f"{k} ({d})"); break
        if names:
            stages[name] = {"kernels": names, "dram_bytes": int(tot)}
    json.dump({"source": source_note, "stages": stages}, open(out, "w"), indent=1)
    print(json.dumps(stages, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
