"""Which fp32 formula does torch's CUDA `tensor / python_float` use? (decides the sampling kernel's rounding)"""
import torch
s = (torch.randn(200000, device="cuda") * 3)
ref = s / 1.05
cands = {
    "fp32 true div by 1.05f": s / torch.tensor(1.05, device="cuda", dtype=torch.float32),
    "mul by fp32(1/1.05f)": s * (torch.tensor(1.0, device="cuda") / torch.tensor(1.05, device="cuda")),
    "double div -> fp32": (s.double() / 1.05).float(),
    "mul by fp32(double 1/1.05)": s * torch.tensor(1.0 / 1.05, device="cuda", dtype=torch.float32),
    "double mul by 1/1.05 -> fp32": (s.double() * (1.0 / 1.05)).float(),
}
for k, v in cands.items():
    print(f"{k:32s} mismatches vs torch: {int((v != ref).sum())}")
m = s * 1.05
print("mul: fp32 x*1.05f mismatches:", int((s * torch.tensor(1.05, device='cuda')) .ne(m).sum()),
      " double:", int((s.double() * 1.05).float().ne(m).sum()))
