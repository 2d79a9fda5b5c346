import re, sys
mx = {}
for line in open(sys.argv[1]):
    m = re.search(r"rel ([0-9.e+-]+)\s*$", line)
    if m:
        k = line.split()[0] + " " + line.split()[1]
        mx[k] = max(mx.get(k, 0), float(m.group(1)))
    if "mismatch" in line or "Error" in line or "Traceback" in line or "diff" in line and "grad" not in line:
        print(line.strip())
for k, v in sorted(mx.items()):
    print(k, v)
