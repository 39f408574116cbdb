"""Development: SASS instruction count per source line of one kernel (needs -lineinfo):
   python tools/sass_lines.py <lib.so> <kernel-name-substring> [top]"""
import collections, re, subprocess, sys, tempfile, os
lib, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout
cnt = collections.Counter(); cur = "?"; active = False
for ln in dis.splitlines():
    m = re.match(r"//-+ \.text\.(\S+) -+", ln)
    if m:
        active = pat in m.group(1); continue
    if not active: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = m.group(1).split("/")[-1] + ":" + m.group(2); continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+[A-Z@]", ln): cnt[cur] += 1
print("total", sum(cnt.values()))
for k, v in cnt.most_common(top): print(f"{v:6d} {k}")
