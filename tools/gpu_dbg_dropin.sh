#!/bin/bash
mkdir -p gpurun_out
S=oracle/_ref/scenes/final.txt
oracle/_ref/rrt_dropin -i $S -w 240 -h 160 -s 8 -o /tmp/a.png 2>/tmp/a.err
rrt_b200/bin/rrt -i $S -w 240 -h 160 -s 8 -o /tmp/b.png 2>/tmp/b.err
rrt_b200/bin/rrt -i $S -w 240 -h 160 -s 8 -o /tmp/c.png 2>/tmp/c.err
python - <<'PY'
import numpy as np
from PIL import Image
a=np.asarray(Image.open('/tmp/a.png')).astype(int); b=np.asarray(Image.open('/tmp/b.png')).astype(int); c=np.asarray(Image.open('/tmp/c.png')).astype(int)
d=np.abs(a-b)
print("differing pixels", (d.max(axis=2)>0).sum(), "max diff", d.max(), "b==c", (b==c).all())
ys,xs=np.nonzero(d.max(axis=2)>0)
print(list(zip(ys[:20],xs[:20])))
for y,x in list(zip(ys,xs))[:10]: print(y,x,a[y,x],b[y,x])
PY
tail -3 /tmp/a.err; tail -3 /tmp/b.err
