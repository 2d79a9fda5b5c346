#!/bin/bash
# development: spills and S2R SR_CgaCtaId sites of k_render<0> by source line (run after make)
cd /tmp && cuobjdump -xelf all /root/repo/scene-aware-3d-multi-human_b200/csrc/mh_render.o >/dev/null 2>&1
nvdisasm --print-line-info mh_render.sm_100a.cubin > render.sass 2>/dev/null
python - <<'PY'
import re
from collections import Counter
cur=None; fn=None; c=Counter(); sp=Counter(); n=0
for line in open('/tmp/render.sass'):
    m=re.search(r'//## File "([^"]+)", line (\d+)',line)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s*\.text\.(\S+):',line)
    if m: fn=m.group(1)
    if fn and 'ILi0' in fn and '/*' in line:
        n+=1
        if 'CgaCtaId' in line: c[cur]+=1
        if re.search(r'\b(STL|LDL)\b',line): sp[cur]+=1
print('instructions', n)
print('CgaCtaId', sorted(c.items(), key=lambda x:x[0][1]))
print('spills', sorted(sp.items(), key=lambda x:x[0][1]))
PY
