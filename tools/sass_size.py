"""SASS instruction counts per function (and per source line for one function) of an object file.
Usage: sass_size.py file.o [function-substring]"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main(obj, sub=None):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
    fn, cur = None, None
    per = collections.defaultdict(collections.Counter)
    for ln in dis.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln) and fn:
            per[fn][cur] += 1
    for f, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
        print(f"{sum(c.values()):6d}  {f[:110]}")
    if sub:
        for f, c in per.items():
            if sub in f:
                print("---", f[:100])
                for (file, l), n in sorted(c.items(), key=lambda kv: -kv[1])[:30]:
                    print(f"{n:5d} {file}:{l}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
