"""Scratch: device vs oracle on a hand-built PP[CON] ALICE desc (first GPU bring-up)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import *
import graniitti_b200 as g
from graniitti_b200 import _abi

def desc_alice(screen):
    d=_abi.ProcessDesc()
    d.abi_version=1; d.ps_class=0; d.proc=15; d.screening=screen; d.ps_active=1; d.lipsdim=6; d.n_tape=4
    d.n_nodes=2; d.n_legs=2; d.n_final=2
    d.loop_nkt=18; d.loop_nphi=15; d.loop_min_kt=1e-4; d.loop_max_kt=1.75
    d.s=169000000.0; d.sqrt_s=13000.0
    d.pbeam1_pz=6499.9999322804224; d.pbeam1_E=6500; d.pbeam2_pz=-6499.9999322804224; d.pbeam2_E=6500
    d.beam1_mass=d.beam2_mass=0.93827208100000004
    d.S_factor=1.0; d.offshell=5.0
    for i,pdg in enumerate([211,-211]):
        d.nodes[i].pdg=pdg; d.nodes[i].parent=-1; d.nodes[i].first_child=-1; d.nodes[i].n_children=0
        d.nodes[i].mass=0.13957061000000001; d.nodes[i].width=2.5283999999999999e-17
    gc=d.gcuts; gc.Y_min=-0.9; gc.Y_max=0.9; gc.M_min=0; gc.M_max=7.0; gc.forward_pt_min=0; gc.forward_pt_max=2.5
    f=d.fcuts; f.active=1; f.eta_min=-0.9; f.eta_max=0.9; f.rap_min=-10; f.rap_max=10; f.pt_min=0.15; f.pt_max=100000; f.Et_min=0; f.Et_max=100000
    f.M_min=0; f.M_max=100000; f.Y_min=-10; f.Y_max=10; f.Pt_min=0; f.Pt_max=100000; f.forward_t_min=0; f.forward_t_max=100000
    s=d.soft; s.DELTA_P=0.098; s.ALPHA_P=0.065; s.gN_P=8.2; s.gN_O=0.1; s.g3P=0.15*8.2; s.fc1=0.34; s.fc2=5.0; s.fc3=10.0
    r=d.regge; r.a0[:]=[1.088,0.55,0.55]; r.ap[:]=[0.2,0.93,0.93]; r.sgn[:]=[1,-1,1]; r.s0=1.0; r.offshellFF=2
    r.b_EXP=0.944; r.a_OREAR=0.638; r.b_OREAR=1.886; r.b_POW=0.5; r.omega=0.5; r.c[:]=[35.00184,81.63672,10.86264]; r.n[:]=[1,0,0]
    return d

msa=np.load(os.path.join(GOLDEN,"msa_pp_13000.npy"))
eik=Eikonal(msa[0],msa[1],msa[2])
print("dfma peak TFLOP/s:", g.measure_dfma_peak(0))
for screen in (0,1):
    d=desc_alice(screen)
    n=200000 if not screen else 50000
    u,tape=make_points(n,6,8,7,adapted=True)
    o=oracle_eval(d,eik if screen else None,u,tape)
    E=g.Engine(desc=d)
    if screen: E.set_eikonal(msa[0],msa[1],msa[2])
    w,fl,dump=E.eval(u.T.copy(),tape.T.copy(),want_dump=True)
    ms,nl=E.last_timing()
    kin=(o['flags']&1)==1
    print("screen",screen,"kernel ms",ms,"launches",nl,"pts/s",n/ms*1e3)
    print(" kin equal",np.array_equal(fl&1,o['flags']&1)," flags equal where kin", np.array_equal(fl[kin],o['flags'][kin]))
    m=o['w']!=0
    print(" zero pattern equal",np.array_equal(w!=0,m), "nonzero",m.sum())
    rel=np.abs(w[m]-o['w'][m])/np.abs(o['w'][m])
    print(" rel err: median %.3g p99 %.3g max %.3g  frac>1e-10: %.3g"%(np.median(rel),np.quantile(rel,0.99),rel.max(),np.mean(rel>1e-10)))
    for k,nm in enumerate(_abi.DUMP_NAMES[:9]):
        a=dump[k][kin]; b=o['dump'][kin,k]
        r=np.abs(a-b)/np.maximum(np.abs(b),1e-300)
        print("   %-6s max rel %.3g"%(nm,r.max()))
    # VEGAS iteration smoke
    xmat=np.tile(((np.arange(128)+1)/128.0)[:,None],(1,6))
    acc=E.vegas_iterate(xmat,0,1<<20,12345678,0)
    ms,nl=E.last_timing()
    print(" vegas: fsum/N=%.6g evals=%d kin=%d fid=%d  ms=%.3f pts/s=%.4g"%(acc.fsum/(1<<20),acc.evaluations,acc.n_kin_ok,acc.n_fid_ok,ms,(1<<20)/ms*1e3))
    E.close()
