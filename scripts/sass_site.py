"""sass_site.py -- prints the SASS of one kernel whose inline chain passes through a given source line.
usage: python scripts/sass_site.py <lib.so> <mangled-kernel-prefix> <file> <line> [--count]"""
import os, re, subprocess, sys, tempfile
lib, kprefix, fname, line = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
for cb in sorted(os.listdir(td)):
    lines = subprocess.run(["nvdisasm", "-gi", os.path.join(td, cb)], capture_output=True, text=True).stdout.splitlines()
    st = [i for i, l in enumerate(lines) if l.startswith(".text." + kprefix)]
    if not st: continue
    en = [i for i in range(st[0] + 1, len(lines)) if lines[i].strip().startswith(".section") or lines[i].startswith(".text.")]
    chain, fresh, out = [], True, []
    for l in lines[st[0]:(en[0] if en else len(lines))]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh: chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', l)
        if m:
            fresh = True
            if any(f == fname and ln == line for f, ln in chain):
                out.append("%-70s [%s]" % (m.group(2).strip(), ">".join(str(c[1]) for c in chain[:3])))
    print(len(out), "instructions")
    if "--count" not in sys.argv: print("\n".join(out))
    break
