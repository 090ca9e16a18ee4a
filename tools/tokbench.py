import sys, os, time, json
sys.path.insert(0, os.getcwd())
import bench, numpy as np
class A: pass
a=A(); a.batch=4096; a.workers=64; a.text_docs=4096; a.text_bytes=2048; a.steps=10
print(json.dumps(bench.text_in_leg(a,0))[:600])
