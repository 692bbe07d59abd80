#!/usr/bin/env python3
"""tools/sass_range.py LIB MANGLED-SUBSTRING START END -- print the SASS of one kernel between two hex offsets"""
import re, subprocess, sys
lib, want, a0, a1 = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", out):
    if want in f.split("\n")[0]:
        for a, t in re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f):
            if a0 <= int(a, 16) <= a1:
                print("%05x  %s" % (int(a, 16), t.strip()))
        break
