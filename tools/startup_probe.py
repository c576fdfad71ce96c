import time, sys, os
sys.path.insert(0, os.getcwd())
t0=time.perf_counter()
import numpy as np
import supersampler_b200 as S
from supersampler_b200 import capi
t1=time.perf_counter(); print(f"import {t1-t0:.3f}")
L=S.device_lib(); H=S.host_lib()
t2=time.perf_counter(); print(f"dlopen {t2-t1:.3f}")
import ctypes as C
n=C.c_int(); L.spsp_device_count(C.byref(n))
t3=time.perf_counter(); print(f"device_count {t3-t2:.3f}")
ctx=S.DeviceContext(31,11,S.threshold(31,11,1000))
t4=time.perf_counter(); print(f"spsp_create {t4-t3:.3f}")
w,nb,ro=S.pack_fasta(b">x\n"+b"ACGT"*1000+b"\n",31)
h=ctx.scan(w,nb)
t5=time.perf_counter(); print(f"first scan (filter build) {t5-t4:.3f}")
h=ctx.scan(w,nb)
t6=time.perf_counter(); print(f"second scan {t6-t5:.3f}")
words,n_,rb,re_,ri=S.batch_layout([w],[ro])
sk=ctx.sketch_batch(words,n_,rb,re_,ri,1,1000)
t7=time.perf_counter(); print(f"first batch (post-pass) {t7-t6:.3f}")
sk=ctx.sketch_batch(words,n_,rb,re_,ri,1,1000)
t8=time.perf_counter(); print(f"second batch {t8-t7:.3f}")
ctx.cmp_load_batch(); ctx.cmp_run((0,1),(0,1),True)
t9=time.perf_counter(); print(f"first compare {t9-t8:.3f}")
