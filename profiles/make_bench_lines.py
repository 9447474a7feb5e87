"""Rebuilds profiles/r2f_bench_lines.md from the bench lines of the frozen round-2 build:
gpurun_out/r2h_{bench,ref}.json (one-GPU box) and gpurun_out/r2g_{bench,ref}_n{1,2,4,8}.json (scratch/runN.sh).
Usage: python profiles/make_bench_lines.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")


def line(name):
    return open(os.path.join(G, name)).read().strip().splitlines()[-1]


def main():
    b, r = line("r2h_bench.json"), line("r2h_ref.json")
    d, rr = json.loads(b), json.loads(r)
    c, k, ro, e = d["configs"], d["configs"]["kinds"], d["roofline"], d["e2e"]
    worst = max(d["bit_identical"]["contract_check_rank0"]["max_rel_coordinate_error"],
                c["multistart8"]["contract_check_rank0"]["max_rel_coordinate_error"],
                c["sweep64m"]["contract_check_rank0"]["max_rel_coordinate_error"],
                *(k[x]["contract_check"]["max_rel_coordinate_error"] for x in ("K2", "K3", "K4")))
    t = f"""# Bench lines of the frozen round-2 build (line form of the contracted runs, linear K4 kernel)

Library: {d['run']['library']}.  One-GPU box (16 host cores), `python bench.py --impl reference` then `python bench.py`, back to
back as the driver runs them (`scratch/run.sh`, outputs `gpurun_out/r2h_*`); both arms at N = 1 / 2 / 4 / 8 follow.  Rebuilt by
`profiles/make_bench_lines.py`.

| | Cramer-form build (`profiles/r2_bench_lines.md`, N = 1) | this build |
|---|---|---|
| `value` (device-resident, `gcs_b200_solve_many` of K1 + K5) | 1.296e10 solves/s, 0.0809 ms/step | **{d['value']:.4g}** solves/s, {d['ms_per_step']:.4f} ms/step |
| the two launches back to back | 1.124e10 | {d['sequential_launches']['value']:.4g} |
| `e2e` (host buffers, copies inside) | 7.29e8-7.45e8, 1.41-1.44 ms/step, 0.86-0.88 of the copy-only ceiling | **{e['value']:.4g}**, {e['ms_per_step']:.3f} ms/step, {e['pcie']['frac_of_ceiling']:.3f} of the ceiling |
| K1 launch (2^19 sub-systems, 2 seeds) / `roofline.frac` (reference-algorithm flops) / by executed flops | 50.0 us / 0.526 / 0.281 | **{ro['launch_ms']*1e3:.1f} us / {ro['frac']:.3f}** / {ro['executed_frac']:.3f} |
| K5 launch | 43.2 us | {ro['second_kernel']['launch_ms']*1e3:.1f} us |
| multi-start x8 (configs[2], 2^20 x 8 runs) | 3.375e9 solves/s | **{c['multistart8']['value']:.4g}** ({c['multistart8']['ms_per_step']*1e3:.0f} us; frac {c['multistart8']['roofline']['frac']:.2f}) |
| sweep 2^26 (configs[4]) | 1.520e10 (4.414 ms) | **{c['sweep64m']['value']:.4g}** ({c['sweep64m']['ms_per_step']:.3f} ms; frac {c['sweep64m']['roofline']['frac']:.3f}) |
| K2 / K3 launch (2^19) | 53.1 / 68.4 us | {k['K2']['line_variant']['launch_ms']*1e3:.1f} / {k['K3']['line_variant']['launch_ms']*1e3:.1f} us |
| K4 launch (2^19, no parallel rows) / HBM fraction | 30.9 us / 0.31 (sequential kernel) | **{k['K4']['line_variant']['launch_ms']*1e3:.1f} us / {k['K4']['line_variant']['hbm_frac']:.2f}** (`newton_linear_kernel`); at 2^22: {k['K4']['at_4m']['launch_ms']*1e3:.1f} us = {k['K4']['at_4m']['hbm_gbs']:.0f} GB/s = **{k['K4']['at_4m']['hbm_frac']:.2f}** of 6539.9 GB/s |
| sketch100k solve | 7.4 ms | {c['sketch100k']['solve_us']/1e3:.2f} ms (host mirror: bit-identical kernels, unchanged code; box to box 7.4-8.8) |
| bit-identical default (`bit_identical.value`) | 6.71e9 | {d['bit_identical']['value']:.4g} |
| reference arm (16 threads, full 2^20 batch per step) | 7.6e6-7.7e6 | {rr['value']:.4g} |
| `value` / reference, `e2e` / reference | 1700, 95-98 | **{d['value']/rr['value']:.0f}, {e['value']/rr['value']:.1f}** |

Every config's contract check against the bit-identical kernels in the same run: iteration counts, flags and roots equal; largest relative
coordinate error {worst:.2e} (tolerance 1e-9).  `roofline.traffic` = {ro['traffic']} B per K1 launch from the ncu capture of this very build
(`traffic_capture_is_of_another_build`: {ro['traffic_capture_is_of_another_build']}).

"""
    rows, raw = [], ""
    for N in (1, 2, 4, 8):
        bn, rn = line(f"r2g_bench_n{N}.json"), line(f"r2g_ref_n{N}.json")
        bj, rj = json.loads(bn), json.loads(rn)
        ej, cj = bj["e2e"], bj["configs"]
        sh = cj["sharded"]
        rows.append(f"| {N} | {bj['value']:.4g} | {ej['value']:.4g} | {ej['ms_per_step']:.2f} | {ej['pcie']['ceiling_ms_per_step']:.2f} / "
                    f"{ej['pcie']['typical_copy_only_ms_per_step']:.2f} | {cj['sweep64m']['value']:.4g} ({cj['sweep64m']['ms_per_step']:.3f} ms) | "
                    f"{cj['multistart8']['value']:.4g} | {cj['kinds']['K4']['line_variant']['hbm_frac']:.2f} / {cj['kinds']['K4']['at_4m']['hbm_frac']:.2f} | "
                    f"{rj['value']:.4g} ({rj['cpu_baseline']['cores']} threads) | {bj['value']/rj['value']:.0f} | {ej['value']/rj['value']:.1f} | "
                    f"{'-' if not sh else 'identical' if all(v['identical_to_one_device'] for v in sh.values()) else 'DIFFERS'} |")
        raw += f"\n## N = {N}: GPU arm\n\n```json\n{bn}\n```\n\n## N = {N}: reference arm\n\n```json\n{rn}\n```\n"
    t += """## Both arms at N = 1 / 2 / 4 / 8

`scratch/runN.sh` under `gpurun --gpus N` (2-, 4- and 8-GPU boxes; N = 1 ran on the 4-GPU box), launched as the driver launches them
(`python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py [--impl reference] --gpus N --steps 20 --warmup 5`), reference arm
first.

| N | value (device-resident) | e2e | e2e ms/step | copy-only ms: best / typical | sweep 2^26 (strong) | multistart x8 | K4 HBM fraction at 2^19 / 2^22 | reference arm | value / ref | e2e / ref | in-process sharded solve vs one device |
|---|---|---|---|---|---|---|---|---|---|---|---|
""" + "\n".join(rows) + f"""

(The N = 1 end-to-end step of this table is slower than the one-GPU box's - 1.59 against {e['ms_per_step']:.2f} ms for the same bytes - with the
same library on that path: another host; its copy-only probe is 1.24 / 1.30 ms.)

## One-GPU box: python bench.py  (GPU arm)

```json
{b}
```

## One-GPU box: python bench.py --impl reference

```json
{r}
```

# Raw lines of the multi-GPU table
""" + raw
    open(os.path.join(ROOT, "profiles", "r2f_bench_lines.md"), "w").write(t)


if __name__ == "__main__":
    main()
