"""Companion of fc1_fp8_experiment.py: HM decision flips (p > 0.5 vs p <= 0.5, Thr_info.txt = 0.5 ...) of FC1 operand
schemes on the reference's 15 000 labelled CTUs, all four deployed QP models, against exact (fp64) arithmetic.  Needs
/root/reference (the CTUs) and the staged checkpoints; CPU only, a few minutes.  Result of round 1 (1.26 M probabilities):
3-pass fp16 0 flips (max |dp| 1.6e-7), fp16 + two fp8 correction passes 2 flips (max |dp| 1.4e-5), plain fp32 0 flips."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import assets  # noqa: E402
from oracle import ethcnn_oracle as eo  # noqa: E402

def f16(v): return v.astype(np.float16).astype(np.float64)
def e4m3(v):
    v=np.asarray(v,np.float64); s=np.sign(v); a=np.minimum(np.abs(v),448.0)
    e=np.maximum(np.floor(np.log2(np.maximum(a,2.0**-20))),-6); step=2.0**(e-3)
    return s*np.round(a/step)*step
def pick(bound): return int(np.floor(np.log2(32768.0/bound)))
files=['AI_Train_5000.dat_shuffled','AI_Valid_5000.dat_shuffled','AI_Test_5000.dat_shuffled']
ctus=np.concatenate([np.fromfile('/root/reference/ETH-CNN_Training_AI/Data/'+f,np.uint8).reshape(-1,4992)[:,:4096].reshape(-1,64,64) for f in files])
print("ctus",ctus.shape)
for qp in (22,27,32,37):
    w=assets.load_weights(assets.AI_MODELS[qp])
    x,q=eo.input_scaling(ctus,qp,eo.MODE_AI,np.float64)
    f=eo.conv_features(x,{k:v.astype(np.float64) for k,v in w.items()})
    W1=np.concatenate([w["h_fc1__%s__w"%h].astype(np.float64) for h in ("64","32","16")],axis=1)
    def heads(z):
        outs=[];o=0
        for h,n1,n2,n3 in eo.HEADS:
            a1=eo._leaky(z[:,o:o+n1]+w["h_fc1__%s__b"%h].astype(np.float64));o+=n1
            a2=eo._leaky(np.concatenate([a1,q],1)@w["h_fc2__%s__w"%h].astype(np.float64)+w["h_fc2__%s__b"%h].astype(np.float64))
            outs.append(eo._sigmoid(np.concatenate([a2,q],1)@w["y_conv_flat__%s__w"%h].astype(np.float64)+w["y_conv_flat__%s__b"%h].astype(np.float64)))
        return np.concatenate(outs,1)
    p_ref=heads(f@W1)
    ef,ew=pick(np.abs(f).max()*1.5),pick(np.abs(W1).max())
    fs,ws=f*2.0**ef,W1*2.0**ew
    fh,wh=f16(fs),f16(ws); fl,wl=f16(fs-fh),f16(ws-wh); un=2.0**-(ef+ew)
    p3=heads((fh@wh+fh@wl+fl@wh)*un)
    p8=heads((fh@wh+(e4m3(fh*2.0**-7)@e4m3(wl*2.0**5))*4.0+(e4m3(fl*2.0**5)@e4m3(wh*2.0**-7))*4.0)*un)
    p32=heads((f.astype(np.float32)@W1.astype(np.float32)).astype(np.float64))
    d=lambda p: (int(((p>0.5)!=(p_ref>0.5)).sum()), float(np.abs(p-p_ref).max()))
    print("qp %d: of %d probabilities  3-pass fp16: flips %d max|dp| %.2g | fp16+2xfp8: flips %d max|dp| %.2g | plain fp32 matmul (the oracle's own arithmetic): flips %d max|dp| %.2g" % ((qp,p_ref.size)+d(p3)+d(p8)+d(p32)))
