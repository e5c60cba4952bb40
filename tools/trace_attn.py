"""Deep trace of the attention phase of the persistent kernel: phase_mask 2 | 32 runs only attention (layer 0..3) with four
extra stamps per item (staged, units done, partial written, counted)."""
import os, sys
os.environ["LIVECC_B200_MEGA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from livecc_b200 import _cabi
from livecc_b200.config import LiveCCConfig
from livecc_b200.engine import LiveCCB200ForConditionalGeneration
cfg = LiveCCConfig.livecc_7b()
eng = LiveCCB200ForConditionalGeneration.from_synthetic(cfg, device="cuda")
kv = int(os.environ.get("KV", "8000"))
g = torch.Generator().manual_seed(0)
out = eng.generate_batch([dict(input_ids=torch.randint(1000, 9000, (1, kv), generator=g).cuda())], max_new_tokens=1)
cache = out[0].past_key_values
with torch.inference_mode():
    cache.scalars[_cabi.SC_FINISHED] = 0
st = [cache.stream_state()]
for _ in range(3):
    eng._native.decode_mega_debug(st, 0, 4, int(os.environ.get("MASK", str(2 | 32))), 0)
torch.cuda.synchronize()
off = eng._native.mega_trace_offset
G = eng.ctx.num_sms
tr = eng._native.workspace[off:off + 256 * 64 * 8].view(torch.int64).view(256, 64)[:G].cpu().double() / 1e3
tr = tr - tr[:, 0].min()
# per CTA per layer: [staged, units, written, counted] (if it had an item), then phase-end stamp, barrier stamp
for cta in list(range(0, 8)) + [40, 90, 127, 147]:
    print(cta, [round(float(x), 2) for x in tr[cta, :16]])
