"""List kernels with register spills / their register counts from csrc/build/*.ptxas.log."""
import glob, re, sys
for path in sorted(glob.glob("wavecap-sdr_b200/csrc/build/*.ptxas.log")):
    name = None
    for line in open(path):
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            name = m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name and (int(m.group(2)) or int(m.group(3)) or "-a" in sys.argv):
            print(path.split("/")[-1].split(".")[0], name[:70], "stack", m.group(1), "st", m.group(2), "ld", m.group(3))
