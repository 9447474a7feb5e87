"""Rebuilds profiles/traffic.json and the round-2 ncu summaries from the .ncu-rep files of one capture
session (gpurun_out/<prefix>_prof_*.ncu-rep, written by scratch/run.sh) and stamps them with the source
hash of the library the capture ran (gpurun_out/<prefix>_ncu_libversion.txt), so that bench.py can tell
whether the library it runs is the build the traffic figures belong to.
Usage: python profiles/make_traffic.py [prefix]      (r2: the Cramer-form build; r2f: the line-form build)"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    units = dict(zip(rows[0], rows[1]))  # ncu picks a unit per column and capture (Kbyte / Mbyte / Gbyte ...)
    return [dict(zip(rows[0], r), _units=units) for r in rows[2:]]


def f(x):
    return float(x.replace(",", ""))


BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def mbytes(d, key):
    return f(d[key]) * BYTES[d["_units"][key]]


def entry(d, src, alg):
    rd, wr = mbytes(d, "dram__bytes_read.sum"), mbytes(d, "dram__bytes_write.sum")
    assert d["_units"]["gpu__time_duration.sum"] in ("us", "usecond"), d["_units"]["gpu__time_duration.sum"]
    cyc = f(d["sm__cycles_elapsed.avg"])
    fl = (f(d["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"])
          + f(d["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"])
          + 2 * f(d["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"])) * cyc
    return {"dram_bytes_per_launch": int((rd + wr) * 1e6), "executed_flops_per_launch": fl, "launch_us_under_ncu": f(d["gpu__time_duration.sum"]),
            "source": f"{src}: dram__bytes_read.sum {rd:.3f} MB + dram__bytes_write.sum {wr:.3f} MB per launch; algorithmic {alg}; "
                      f"executed FP64 flops (DADD + DMUL + 2 DFMA) {fl:.3e} per launch"}


def summary(rep, title, dst):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "summarize_ncu.py"), rep, title], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", dst), "w").write(txt)


def main():
    P = sys.argv[1] if len(sys.argv) > 1 else "r2f"
    lib = open(os.path.join(OUT, f"{P}_ncu_libversion.txt")).read().strip()
    h = re.search(r"src ([0-9a-f]+)", lib).group(1)
    t = {"_src_hash": h, "_library": lib,
         "_note": f"per-launch DRAM bytes and executed FP64 flops from the round-2 `ncu --set full` captures (profiles/{P}_ncu_*.md); bench.py "
                  "reports `traffic_capture_is_of_another_build` when the library it runs carries another source hash"}
    st = raw(os.path.join(OUT, f"{P}_prof_static.ncu-rep"))
    t["newton_static_kernel[contracted]<K1,2 seeds>"] = entry(st[0], f"profiles/{P}_ncu_contracted.md", "37.7 MB (the 12 MB of outputs are still in the 126 MB L2 when the kernel ends)")
    t["newton_static_kernel[contracted]<K5,2 seeds>"] = entry(st[1], f"profiles/{P}_ncu_contracted.md", "75.5 MB")
    so = raw(os.path.join(OUT, f"{P}_prof_sorted.ncu-rep"))
    t["newton_sorted_kernel<K1,2 seeds>"] = entry(so[0], f"profiles/{P}_ncu_sorted.md", "37.7 MB")
    t["newton_sorted_kernel<K5,2 seeds>"] = entry(so[1], f"profiles/{P}_ncu_sorted.md", "75.5 MB")
    k4 = raw(os.path.join(OUT, f"{P}_prof_seq_k4.ncu-rep"))
    name = "newton_seq_kernel<K4,2 seeds>" if "false" in k4[0]["Kernel Name"] or ", 0>" in k4[0]["Kernel Name"] else "newton_seq_kernel[contracted]<K4,2 seeds>"
    t[name] = entry(k4[0], f"profiles/{P}_ncu_seq_k4.md", "62.9 MB")
    lin = os.path.join(OUT, f"{P}_prof_linear_k4.ncu-rep")
    if os.path.exists(lin):
        t["newton_linear_kernel[contracted]<K4,2 seeds>"] = entry(raw(lin)[0], f"profiles/{P}_ncu_linear_k4.md", "62.9 MB")
    k8 = raw(os.path.join(OUT, f"{P}_prof_seq_k1x8.ncu-rep"))
    t["newton_seq_kernel[contracted]<K1,8 seeds>"] = entry(k8[0], f"profiles/{P}_ncu_seq_k1x8.md", "2^20 x (48 + 1 + 16 + 1 + 24) B = 94.4 MB")
    json.dump(t, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    cmd = "ncu --set full --clock-control none --import-source on"
    summary(os.path.join(OUT, f"{P}_prof_static.ncu-rep"),
            f"Contracted static kernels (GCS_VARIANT_CONTRACTED for K1 / K5 with 2 seeds = newton_static_kernel<KIND, 2, true>), round 2\n\nlibrary: {lib}\n\n"
            f"command: `{cmd} -k regex:newton_static_kernel -s 10 -c 2 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline` (after the same "
            f"command exited 0 without ncu); launch list of the same command: `profiles/{P}_launches.csv`", f"{P}_ncu_contracted.md")
    summary(os.path.join(OUT, f"{P}_prof_sorted.ncu-rep"),
            f"Bit-identical default kernels (GCS_VARIANT_DEFAULT at 2^19 = newton_sorted_kernel<KIND, 2, 128, 128, false>), round 2\n\nlibrary: {lib}\n\n"
            f"command: `{cmd} -k regex:newton_sorted_kernel -s 4 -c 2 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline`", f"{P}_ncu_sorted.md")
    summary(os.path.join(OUT, f"{P}_prof_seq_k4.ncu-rep"),
            f"Sequential kernel on K4 (newton_seq_kernel<4, 2, .>, 2^19 sub-systems without parallel line pairs), round 2\n\nlibrary: {lib}\n\n"
            f"command: `{cmd} -k regex:newton_seq_kernel -s 4 -c 1 python scratch/k4_hbm.py`", f"{P}_ncu_seq_k4.md")
    if os.path.exists(lin):
        summary(lin, f"Linear kernel on K4 (newton_linear_kernel<2>: GCS_VARIANT_CONTRACTED for K4; 2^19 sub-systems without parallel line pairs), round 2\n\nlibrary: {lib}\n\n"
                f"command: `{cmd} -k regex:newton_linear_kernel -s 4 -c 1 python scratch/k4_hbm.py`", f"{P}_ncu_linear_k4.md")
    summary(os.path.join(OUT, f"{P}_prof_seq_k1x8.ncu-rep"),
            f"Sequential kernel on the 8-seed K1 (newton_seq_kernel<1, 8, true>, 2^20 sub-systems x 8 seeds: configs[2]), round 2\n\nlibrary: {lib}\n\n"
            f"command: `{cmd} -k regex:newton_seq_kernel -s 3 -c 1 python scratch/kbench.py 5 1 1048576 8`", f"{P}_ncu_seq_k1x8.md")
    for k, v in t.items():
        if isinstance(v, dict):
            print(k, v["dram_bytes_per_launch"], f"{v['executed_flops_per_launch']:.3e}", v["launch_us_under_ncu"])


if __name__ == "__main__":
    main()
