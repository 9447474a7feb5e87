"""Loop statistics from SASS: for every backward branch of a kernel, the loop's instruction count,
FP64 count and issue cycles (an FP64 instruction holds a scheduler's issue port for two cycles on
this part, any other for one: DESIGN.md section 3), and - with --sass - the instructions of the
innermost (smallest) loops, i.e. the update loop the kernel lives in.
Usage: python profiles/loopstat.py lib.so kernel-substring [--sass N]"""
import re
import subprocess
import sys

FP64 = ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, funcs = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2)))
    return funcs


def is_fp64(text):
    toks = text.replace("@", " ").split()
    return any(t.startswith(FP64) for t in toks[:2])


def loops(ins):
    addr = {a: i for i, (a, _) in enumerate(ins)}
    found = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr:
            j = addr[int(m.group(1), 16)]
            body = ins[j:i + 1]
            nf = sum(1 for _, x in body if is_fp64(x))
            mufu = sum(1 for _, x in body if "MUFU" in x)
            found.append((j, i, len(body), nf, mufu, 2 * nf + len(body) - nf))
    return found


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    n_sass = int(sys.argv[sys.argv.index("--sass") + 1]) if "--sass" in sys.argv else 0
    for name, ins in functions(lib).items():
        if pat not in name:
            continue
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        print(f"### `{demangled}` ({len(ins)} SASS instructions)\n")
        ls = loops(ins)
        print("| loop (instruction index) | instructions | FP64 | MUFU | issue cycles = 2 x FP64 + other |\n|---|---|---|---|---|")
        for j, i, n, nf, mufu, cyc in ls:
            print(f"| {j}-{i} | {n} | {nf} | {mufu} | {cyc} |")
        print()
        for j, i, n, nf, mufu, cyc in sorted((l for l in ls if l[2] > 8), key=lambda l: l[2])[:n_sass]:
            print(f"innermost loop {j}-{i} ({n} instructions, {nf} FP64):\n\n```")
            for a, t in ins[j:i + 1]:
                print(f"/*{a:04x}*/ {t}")
            print("```\n")


if __name__ == "__main__":
    main()
