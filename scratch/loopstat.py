"""Loop statistics from SASS: for every backward branch, the loop's instruction count, FP64 count and
issue cycles (2*FP64 + other).  Usage: python scratch/loopstat.py lib.so kernel-substring"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2)))
FP64 = ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")
for name, ins in funcs.items():
    if pat not in name:
        continue
    print(name, len(ins), "instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr:
            j = addr[int(m.group(1), 16)]
            body = ins[j:i + 1]
            nf = sum(1 for _, x in body if x.split()[0].lstrip("@!P0123456789T ").startswith(FP64) or any(x.replace("@", " ").split()[k].startswith(FP64) for k in range(min(2, len(x.split())))))
            mufu = sum(1 for _, x in body if "MUFU" in x)
            print(f"  loop {j}-{i}: {len(body)} instr, {nf} fp64, {mufu} mufu, issue cycles {2*nf + len(body)-nf}")
