"""Failure rate of the mode-0 Taylor series (csrc/rbk_math.cuh, compiled for the host) by order and time step, for TIP3P water at
300 K: the numbers behind the series ladder {11, 13, 16}.  python tools/series_stats.py"""
import ctypes as C, numpy as np, sys
sys.path.insert(0,'/root/repo/tests')
lib=C.CDLL('/root/repo/openmm_rigidbody_plugin_b200/lib/librbk_hostmath.so')
_dp=C.POINTER(C.c_double)
lib.rbkh_exact_series.argtypes=[C.c_int,C.c_double,_dp,_dp,_dp]; lib.rbkh_exact_series.restype=C.c_int
rng=np.random.Generator(np.random.Philox(key=5))
# actual TIP3P principal moments
import numpy as np
from openmm_rigidbody_plugin_b200 import synth
def water_I():
    half=0.5*synth.ANGLE_HOH
    site=np.array([[0,0,0],[synth.R_OH*np.sin(half),0,synth.R_OH*np.cos(half)],[-synth.R_OH*np.sin(half),0,synth.R_OH*np.cos(half)]])
    m=np.array([synth.M_O,synth.M_H,synth.M_H]); c=(m[:,None]*site).sum(0)/m.sum(); d=site-c
    T=sum(mi*(np.dot(x,x)*np.eye(3)-np.outer(x,x)) for mi,x in zip(m,d))
    return np.sort(np.linalg.eigvalsh(T))[::-1].copy()
I=water_I(); print("I",I)
N=100000
samples=[]
for _ in range(N):
    Lb=rng.standard_normal(3)*np.sqrt(synth.KT_300K*I)
    q=rng.standard_normal(4); q/=np.linalg.norm(q)
    pi=2*np.array([-q[1]*Lb[0]-q[2]*Lb[1]-q[3]*Lb[2], q[0]*Lb[0]-q[3]*Lb[1]+q[2]*Lb[2], q[3]*Lb[0]+q[0]*Lb[1]-q[1]*Lb[2], -q[2]*Lb[0]+q[1]*Lb[1]+q[0]*Lb[2]])
    samples.append((q,pi))
def frac(order, dt):
    bad=0
    for q,pi in samples:
        qq=q.copy(); pp=pi.copy()
        ok=lib.rbkh_exact_series(order, dt, I.ctypes.data_as(_dp), qq.ctypes.data_as(_dp), pp.ctypes.data_as(_dp))
        bad+= ok!=1
    return bad/N
for dt in (0.001, 0.002, 0.003, 0.004, 0.005):
    print(dt, {k: frac(k,dt) for k in (10, 11, 12, 13, 14, 16)})
