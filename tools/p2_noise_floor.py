"""TEST INFRASTRUCTURE.  How far apart are two CPU runs of the reference's arithmetic after 1 / 2 / 5 free-running guided steps with the
shipped (trained) IS checkpoints?  The fp32 restatement with 1 and 8 threads and an fp64 run of it against the fp32 reference's stored
latents (tests/golden/p4_is_small.npz).  Output committed as profiles/r02_p2_noise_floor_is.txt: it is the yardstick for the P2 numbers
of the GPU run (profiles/r02_p4_report.json) -- the trajectory amplifies rounding noise by ~400x per step, and one of the four instances
is O(0.1) away from ITSELF after five steps."""
import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/diffusion-integer-programming_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from dip_b200 import synth
from oracle import restate
from oracle.gen_golden2 import p4_cases, P4
from helpers import golden, tw, rel_l2
kind="is"
g=golden("p4_%s_small"%kind); cfg=P4[kind]
Wd=torch.load('/root/repo/tests/golden/ddpm_independent_set.pth',weights_only=True, map_location='cpu'); Wdec=torch.load('/root/repo/tests/golden/decoder_independent_set.pth',weights_only=True, map_location='cpu')
for i,(inst,mip,kpm,L) in enumerate(p4_cases(kind)):
    for dt,nt in ((torch.float32,1),(torch.float32,8),(torch.float64,8)):
        torch.set_num_threads(nt)
        B=8
        x=torch.from_numpy(synth.initial_noise(cfg["inst_base"]+i,range(B),L)).to(dt)
        mipB,kpmB=torch.from_numpy(mip)[None].repeat(B,1,1).to(dt),torch.from_numpy(kpm)[None].repeat(B,1)
        A=(torch.from_numpy(inst.rows),torch.from_numpy(inst.cols),torch.from_numpy(inst.a_val).to(dt),inst.M)
        tab=restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"],cfg["S"])
        W1={k:v.to(dt) for k,v in Wd.items()}; W2={k:v.to(dt) for k,v in Wdec.items()}
        out=[]
        for k in range(5):
            x=restate.guided_ddim_step(W1,W2,x,cfg["S"]-1-k,restate.ddim_step_coeffs(tab,k),mipB,kpmB,A,torch.from_numpy(inst.b).to(dt),torch.from_numpy(inst.c).to(dt),cfg["gs"],cfg["lam"])["x_prev"]
            if k+1 in (1,2,5): out.append(float("%.2g"%rel_l2(x[:,:inst.n],g["i%d_x_after%d"%(i,k+1)][:B,:inst.n])))
        print("inst",i, dt, "threads",nt, out, flush=True)
