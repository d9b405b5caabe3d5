"""sass_sizes.py -- code bytes of a kernel and of the out-of-line device functions inside it (from the ELF symbol table).
usage: python tools/sass_sizes.py lib.so kernel_name_substring"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-elf", sys.argv[1]], capture_output=True, text=True).stdout
pat = sys.argv[2]
rows = []
for ln in out.splitlines():
    m = re.match(r"\s+0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+\S+\s+(\S+)$", ln)
    if m and pat in m.group(3):
        name = m.group(3).split("$")[-1] or m.group(3)
        rows.append((int(m.group(1), 16), int(m.group(2), 16), name))
for ln in out.splitlines():
    m = re.match(r"\s+[0-9a-f]+\s+[0-9a-f]+\s+([0-9a-f]+)\s+\S+\s+\S+\s+PROGBITS.*\.text\.(\S+)$", ln)
    if m and pat in m.group(2):
        print("section .text.%s: %d bytes" % (m.group(2)[:60], int(m.group(1), 16)))
for off, size, name in sorted(set(rows)):
    print("  +0x%05x %6d  %s" % (off, size, name[:90]))
