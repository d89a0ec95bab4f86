import numpy as np, sys
sys.path.insert(0,'/root/repo')
from ferromic_b200 import _lib
L=_lib.lib()
rng=np.random.default_rng(6)
n=1<<22
b=np.concatenate([rng.random(n)*0.999+0.001, np.array([1.0,0.5,0.25,np.nextafter(1.0,0.0),np.nextafter(0.5,1.0),0.75,1.0-2.0**-30])])
a=b.copy(); y=np.empty_like(b); q=np.empty_like(b)
_lib.check(L.fm_wc_arith_probe(a.ctypes.data,b.ctypes.data,y.ctypes.data,q.ctypes.data,b.size))
ref=1.0/b
bad=np.nonzero(y.view(np.uint64)!=ref.view(np.uint64))[0]
print("mismatches",bad.size,"of",b.size)
for i in bad[:10]:
    print(float(b[i]).hex(), float(y[i]).hex(), float(ref[i]).hex(), int(y[i].view(np.uint64))-int(ref[i].view(np.uint64)))
