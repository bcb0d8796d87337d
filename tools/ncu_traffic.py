#!/usr/bin/env python
"""Turns an `ncu --set full --page raw --csv` export of one bench.py step into profiles/r2_dram_traffic.json, the file
bench.py reads `roofline.traffic` from (dram__bytes_read.sum + dram__bytes_write.sum per launch, at the benched launch
size), plus a launch list with durations and the pipe utilisations of every kernel.

  # on the GPU box (one step of the default workload after the warm-up; ~40 replays per kernel):
  ncu --set full --clock-control none -k regex:"k_phi_fft|k_fwd_uni|k_fwd_pipe|k_legendre|k_dct|k_inv_uni" -s 15 -c 5 \
      -o gpurun_out/step python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --single-bw 0 --no-strong
  # here:
  ncu -i gpurun_out/step.ncu-rep --page raw --csv > /tmp/step.csv; python tools/ncu_traffic.py /tmp/step.csv 256 1024
"""
import csv
import re
import json
import os
import sys

KIND = {"k_phi_fft_fwd": "phi_fft_fwd", "k_phi_fft_inv": "phi_fft_inv", "k_fwd_uni": "fused_fwd", "k_fwd_pipe": "fused_fwd",
        "k_inv_uni": "fused_inv", "k_legendre_fwd": "legendre_fwd", "k_leg_fwd_stream": "legendre_fwd",
        "k_legendre_inv": "legendre_inv", "k_dct_fwd": "dct_fwd", "k_dct_inv": "dct_inv"}
GB = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    path, bw, nfun = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=None):
        i = col[name]
        v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0.0
        return v * (GB.get(units[i], 1.0) if scale == "bytes" else 1.0)

    kernels, launches = {}, []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        kind = next((v for k, v in KIND.items() if re.search(r"(^|[\s:])" + k, name)), None)
        if kind is None:
            continue
        dram = val(r, "dram__bytes_read.sum", "bytes") + val(r, "dram__bytes_write.sum", "bytes")
        ent = {"kernel": name.split("(")[0].replace("void ", ""), "functions_per_launch": nfun, "dram_bytes_per_launch": dram,
               "dram_read_bytes": val(r, "dram__bytes_read.sum", "bytes"), "dram_write_bytes": val(r, "dram__bytes_write.sum", "bytes"),
               "duration_under_ncu": val(r, "gpu__time_duration.sum"), "duration_unit": units[col["gpu__time_duration.sum"]],
               "fp64_pipe_pct": val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
               "dmma_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
               "lsu_wavefronts_pct": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
               "registers_per_thread": val(r, "launch__registers_per_thread")}
        kernels[kind] = ent
        launches.append(ent)
    out = {"bw": bw, "format_real": False, "source": "ncu --set full --clock-control none, one step of `python bench.py` "
           "(bw 256, 1024 functions per launch, COMPLEX) after the warm-up; durations are cold-cache and serialised",
           "kernels": kernels}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_dram_traffic.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
