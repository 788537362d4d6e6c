import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import gorender_b200 as g
from gorender_b200 import parallel, workloads
W, H = 3840, 2160
objs, cam = workloads.config_c4(100)
dev = g.default_device(0)
fb = g.FrameBuffer(W, H, 1, dev)
r = g.Renderer(fb)
packed = np.ascontiguousarray(r.pack_objects(objs, [cam]))
for rows in (None, (0, 1088), (544, 832), (1088, 1248)):
    r.draw_packed(packed, 0, rows=rows)
    print(rows, int(r.last_stats["triangles"][0]), int(r.last_stats["tpf"][0]))
