"""SASS instructions per source region of one kernel: `nvdisasm -g -c` output of the library's cubin, sliced to the
kernel, instructions attributed to the (file, line) of the preceding line-info comment and summed per bucket of lines.
Usage: python tools/sass_by_line.py [kernel-substring] [bucket]   (default: the config-2 beam kernel, 25 lines)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("CTCDECODE_B200_LIB") or os.path.join(ROOT, "ctcdecode_b200", "_lib", "libctcdecode_b200.so")


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else "beam_kernelILi256ELb0ELb0ELi128ELb0E"
    bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, check=True, stdout=subprocess.PIPE,
                              stderr=subprocess.DEVNULL).stdout.decode()
    inside, cur, n = False, None, 0
    per = collections.Counter()
    for line in text.splitlines():
        if line.startswith(".text."):
            inside = want in line
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            per[(cur[0], cur[1] // bucket * bucket)] += 1
            n += 1
    print("kernel *%s*: %d SASS instructions (%d KB)" % (want, n, n * 16 // 1024))
    for (f, l), c in sorted(per.items()):
        print("%-34s %5d-%-5d %5d" % (f, l, l + bucket - 1, c))


if __name__ == "__main__":
    main()
