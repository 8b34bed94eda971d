"""Static opcode histogram of the loops of a kernel in a built library or cubin (no GPU, no ncu): every backward branch
closes a loop; loops shorter than `min_len` instructions are skipped.
usage: python profiles/sass_loops.py fdtd-2d_b200/libfdtd2d.so <mangled-name-substring> [min_len=300]
e.g.   python profiles/sass_loops.py fdtd-2d_b200/libfdtd2d.so strip_wave_x2_kernelILi8ELb1ELi3ELb1ELi0E
       (the third loop listed is the plain-strip loop, two rows per trip: DESIGN.md section 9)"""
import collections
import re
import subprocess
import sys

binary, pattern = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 300
out = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
for f in re.split(r"\n\s+Function : ", out)[1:]:
    name = f.split("\n")[0]
    if pattern not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(f"{name}: {len(ins)} instructions")
    for addr, text in ins:
        m = re.search(r"BRA\S*\s+(?:.*?)0x([0-9a-f]+)", text)
        if not m or int(m.group(1), 16) >= addr:
            continue
        target = int(m.group(1), 16)
        body = [t for a, t in ins if target <= a <= addr]
        if len(body) < min_len:
            continue
        ops = collections.Counter(re.sub(r"@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
        print(f"  loop {target:#x}..{addr:#x}: {len(body)} instructions: " + "  ".join(f"{k} {v}" for k, v in ops.most_common(12)))
