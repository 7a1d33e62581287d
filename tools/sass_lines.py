"""Code size of one kernel per source range (instruction-cache budget): disassembles the built library with line
information and counts SASS instructions per marked range of the kernel source.
Usage: python tools/sass_lines.py ac3_encode.cu ac3_encode_kernel"""
import collections, os, re, subprocess, sys, tempfile
src_name, kern = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "ac-3-acm-codec_b200", "liba52_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cnt = collections.Counter()
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn, cur = False, None
    for l in out.split("\n"):
        if ".text." in l:
            infn = kern in l
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if infn and re.match(r"\s+/\*[0-9a-f]+\*/\s+[A-Z@]", l):
            cnt[cur] += 1
tot = sum(cnt.values())
print("total", tot, "instructions,", tot * 16 // 1024, "KB")
src = open(os.path.join(root, "ac-3-acm-codec_b200", "csrc", src_name)).read().split("\n")
marks = [(i, l.strip()[:70]) for i, l in enumerate(src, 1)
         if "=================" in l or l.startswith("__device__") or l.startswith("__global__") or l.startswith("template <")]
marks.append((len(src) + 1, "end"))
for (a, n), (b, _) in zip(marks, marks[1:]):
    s = sum(v for (f, ln), v in cnt.items() if f == src_name and a <= ln < b)
    if s:
        print("%5d-%5d %6d  %s" % (a, b - 1, s, n))
print("other files", sum(v for (f, ln), v in cnt.items() if f != src_name))
