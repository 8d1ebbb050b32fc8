"""Numpy experiment behind DESIGN.md section 3.2 ("what would move it next"): how much probability error do cheaper
operand schemes for the FC1 contraction cost?  FC2 / FC3 / sigmoid are evaluated exactly (fp64) so that only FC1's operand
rounding shows.  Schemes: one fp16 pass; fp16 A with exact W; today's three fp16 passes (hi*hi + hi*lo + lo*hi); one fp16
pass + two fp8 (e4m3) correction passes.  CPU only; uses the deployed QP-32 checkpoint and, when /root/reference is
present, 1500 of the reference's labelled CTUs next to procedural frames."""
import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import ethcnn_oracle as eo, assets
w = assets.load_weights(assets.AI_MODELS[32])
rng = np.random.default_rng(0)
frames = [eo.synth_frame(1920,1080,s) for s in range(3)]
ctus = np.concatenate([eo.frame_to_ctus(np.pad(f,((0,8),(0,0)))) for f in frames])[:1200]
try:
    raw = np.fromfile('/root/reference/ETH-CNN_Training_AI/Data/AI_Test_5000.dat_shuffled', np.uint8).reshape(-1,4992)[:1500,:4096].reshape(-1,64,64)
    ctus = np.concatenate([ctus, raw])
except Exception as e: print(e)
x,q = eo.input_scaling(ctus, 32, eo.MODE_AI, np.float64)
f = eo.conv_features(x, {k:v.astype(np.float64) for k,v in w.items()})
print("ctus", ctus.shape, "feat max", np.abs(f).max())
W1 = np.concatenate([w["h_fc1__%s__w"%h].astype(np.float64) for h in ("64","32","16")],axis=1)
def heads_from_a1pre(z):
    outs=[]; o=0
    for h,n1,n2,n3 in eo.HEADS:
        a1 = eo._leaky(z[:,o:o+n1] + w["h_fc1__%s__b"%h].astype(np.float64)); o+=n1
        a2 = eo._leaky(np.concatenate([a1,q],1) @ w["h_fc2__%s__w"%h].astype(np.float64) + w["h_fc2__%s__b"%h].astype(np.float64))
        outs.append(eo._sigmoid(np.concatenate([a2,q],1) @ w["y_conv_flat__%s__w"%h].astype(np.float64) + w["y_conv_flat__%s__b"%h].astype(np.float64)))
    return np.concatenate(outs,1)
p_ref = heads_from_a1pre(f @ W1)
def f16(v): return v.astype(np.float16).astype(np.float64)
def e4m3(v):
    v=np.asarray(v,np.float64); s=np.sign(v); a=np.abs(v)
    a=np.minimum(a,448.0)
    e=np.floor(np.log2(np.maximum(a,2.0**-20))); e=np.maximum(e,-6)
    step=2.0**(e-3)
    return s*np.round(a/step)*step
def pick(bound): return int(np.floor(np.log2(32768.0/bound)))
ef, ew = pick(np.abs(f).max()*1.5), pick(np.abs(W1).max())
fs, ws = f*2.0**ef, W1*2.0**ew
fh, wh = f16(fs), f16(ws); fl, wl = f16(fs-fh), f16(ws-wh)
un = 2.0**-(ef+ew)
z3 = (fh@wh + fh@wl + fl@wh)*un
print("3-pass fp16:      max|dp| = %.3g" % np.abs(heads_from_a1pre(z3)-p_ref).max())
print("1-pass fp16:      max|dp| = %.3g" % np.abs(heads_from_a1pre((fh@wh)*un)-p_ref).max())
print("2-pass (w exact): max|dp| = %.3g" % np.abs(heads_from_a1pre((fh@wh+fh@wl)*un)-p_ref).max())
for sh_hi in (7,8):
  for sh_lo in (4,5,6):
    fh8 = e4m3(fh*2.0**-sh_hi); wh8 = e4m3(wh*2.0**-sh_hi)
    fl8 = e4m3(fl*2.0**sh_lo); wl8 = e4m3(wl*2.0**sh_lo)
    z8 = (fh@wh + (fh8@wl8)*2.0**(sh_hi-sh_lo) + (fl8@wh8)*2.0**(sh_hi-sh_lo))*un
    print("fp16 + 2x fp8 corrections (hi>>%d, lo<<%d): max|dp| = %.3g   lo8 max %.3g hi8 max %.3g" % (sh_hi, sh_lo, np.abs(heads_from_a1pre(z8)-p_ref).max(), np.abs(fl*2.0**sh_lo).max(), np.abs(fh*2.0**-sh_hi).max()))
