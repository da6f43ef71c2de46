"""Where does the Cnn14 encoder's deviation from the oracle come from?  body on the oracle's log-mel vs end to end."""
import os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from audiocaption_b200 import _lib
from audiocaption_b200.captioning.models.cnn_encoder import Cnn14Encoder
from oracle import cnn14 as oc, caption_model as cm
dev = "cuda:0"
sd = oc.build_state_dict(3)
m = Cnn14Encoder().eval(); m.load_state_dict(sd, strict=True); m = m.to(dev)
wav, lens = cm.synth_wav(3, 64000, seed=9, ragged=True, varied=True, sample_rate=32000)
lms = oc.log_mel(sd, wav)
fl = oc.feat_lengths(lens)
ref_attn, ref_fc = oc.body(sd, lms, fl)
l = _lib.lib()
B, F, T = lms.shape
Tp = l.ac_cnn14_out_frames(T)
attn = torch.empty(B, Tp, 2048, device=dev); fc = torch.empty(B, 2048, device=dev)
nb = l.ac_cnn14_workspace_bytes(B, F, T)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
lms_d = lms.to(dev).contiguous(); len_d = fl.to(dev)
_lib.check(l.ac_cnn14_fwd(m._net(), _lib.ptr(lms_d), B, F, T, _lib.ptr(len_d), _lib.ptr(attn), _lib.ptr(fc), _lib.ptr(ws), nb,
                          _lib.current_stream()), "fwd")
torch.cuda.synchronize()
print("body on oracle lms: attn rel", ((attn.cpu() - ref_attn).abs().max() / ref_attn.abs().max()).item(),
      "fc rel", ((fc.cpu() - ref_fc).abs().max() / ref_fc.abs().max()).item())
glms = m.log_mel(wav.to(dev)).cpu()
d = (glms - lms).abs()
print("log-mel: max dB err", d.max().item(), "mean", d.mean().item(), "lms range", lms.min().item(), lms.max().item())
i = d.argmax(); print("worst at", np.unravel_index(i.item(), d.shape), "ref", lms.flatten()[i].item())
out = m({"wav": wav.to(dev), "wav_len": lens, "specaug": False})
print("end to end: attn rel", ((out["attn_emb"].cpu() - ref_attn).abs().max() / ref_attn.abs().max()).item())
# sensitivity of the ORACLE itself to a 1e-3 dB perturbation of its input
pert = lms + 1e-3 * torch.randn(lms.shape, generator=torch.Generator().manual_seed(0))
pa, _ = oc.body(sd, pert, fl)
print("oracle sensitivity to 1e-3 dB noise: attn rel", ((pa - ref_attn).abs().max() / ref_attn.abs().max()).item())
