import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from walkgpt_b200.modules import PromptEncoder, MaskDecoderMultiScale
from oracle import path_a
dev = "cuda"
torch.manual_seed(0)

def cmp(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs().max().item()
    print(f"{name}: max_abs_err={err:.4e} ref_absmax={ref.abs().max().item():.4e} shape={tuple(ref.shape)}")

g, S = 8, 3
hw = g * g
pe_m = PromptEncoder(256, (g, g), (g * 14, g * 14), 16, seed=6).cuda()
dec = MaskDecoderMultiScale(seed=7).cuda()
emb = torch.randn(1, 256, g, g, device=dev)
txt = torch.randn(S, 1, 256, device=dev) * 0.5
sparse, dense = pe_m(None, None, None, txt)
pe = pe_m.get_dense_pe()

sdd = {k: v.detach().float().cpu() for k, v in dec.state_dict().items()}
def attn(p, q, k, v):
    return path_a._sam_attn(sdd, p, q, k, v, 8, path_a.FP32)
def ln(x, n):
    return F.layer_norm(x, (256,), sdd[n + ".weight"], sdd[n + ".bias"])
out_tok = torch.cat([sdd["iou_token.weight"], sdd["mask_tokens.weight"]], 0)
tokens = torch.cat([out_tok[None].expand(S, -1, -1), sparse.cpu()], 1) + sdd["level_embed.weight"][0][None, None]
T = "transformer.0."
lp = T + "layers.0."
keys = (emb.bfloat16().float().cpu() + dense.cpu()[:1]).flatten(2).permute(0, 2, 1).expand(S, -1, -1)
kpe = pe.cpu().flatten(2).permute(0, 2, 1).expand(S, -1, -1)
P = S; rows = P * hw
def views():
    global off
    off = 0
    ws = dec._ws.buf
    def take(nbytes, dtype, shape):
        global off
        a = (off + 255) & ~255
        off = a + nbytes
        return ws[a:a + nbytes].view(dtype).reshape(shape)
    v = {}
    v["keysA"] = take(rows * 256 * 2, torch.bfloat16, (P, hw, 256))
    v["keysB"] = take(rows * 256 * 2, torch.bfloat16, (P, hw, 256))
    v["kvq"] = take(rows * 384 * 2, torch.bfloat16, (P, hw, 384))
    v["a2"] = take(rows * 128 * 2, torch.bfloat16, (P, hw, 128))
    v["U"] = take(rows * 128 * 4, torch.float32, (P, hw, 4, 32))
    v["Tq"] = take(P * 6 * 256 * 4, torch.float32, (P, 6, 256))
    v["Tpe"] = take(P * 6 * 256 * 4, torch.float32, (P, 6, 256))
    v["KT"] = take(P * 6 * 128 * 4, torch.float32, (P, 6, 128))
    v["VT"] = take(P * 6 * 128 * 4, torch.float32, (P, 6, 128))
    return v
W = lambda n: sdd[lp + n + ".weight"]
Bi = lambda n: sdd[lp + n + ".bias"]
for stop in (1, 2, 3, 4):
    os.environ["WG_DEBUG_DECODER_STOP"] = str(stop)
    dec(emb, pe, sparse, dense, True, 0)
    torch.cuda.synchronize()
    v = views()
    if stop == 1:
        cmp("keys0", v["keysA"], keys)
        kk = F.linear(keys + kpe, W("cross_attn_token_to_image.k_proj"), Bi("cross_attn_token_to_image.k_proj"))
        vv = F.linear(keys, W("cross_attn_token_to_image.v_proj"), Bi("cross_attn_token_to_image.v_proj"))
        qq = F.linear(keys + kpe, W("cross_attn_image_to_token.q_proj"), Bi("cross_attn_image_to_token.q_proj"))
        cmp("kvq.K", v["kvq"][..., :128], kk); cmp("kvq.V", v["kvq"][..., 128:256], vv); cmp("kvq.Q", v["kvq"][..., 256:], qq)
    if stop == 2:
        q = attn(lp + "self_attn.", tokens, tokens, tokens)
        q = ln(q, lp + "norm1")
        q = ln(q + attn(lp + "cross_attn_token_to_image.", q + tokens, keys + kpe, keys), lp + "norm2")
        h = torch.relu(F.linear(q, W("mlp.lin1"), Bi("mlp.lin1")))
        q = ln(q + F.linear(h, W("mlp.lin2"), Bi("mlp.lin2")), lp + "norm3")
        kt = F.linear(q + tokens, W("cross_attn_image_to_token.k_proj"), Bi("cross_attn_image_to_token.k_proj"))
        vt = F.linear(q, W("cross_attn_image_to_token.v_proj"), Bi("cross_attn_image_to_token.v_proj"))
        cmp("Tq after layer0", v["Tq"], q); cmp("KT", v["KT"], kt); cmp("VT", v["VT"], vt)
    if stop == 3:
        qi = F.linear(keys + kpe, W("cross_attn_image_to_token.q_proj"), Bi("cross_attn_image_to_token.q_proj")).view(S, hw, 8, 16).transpose(1, 2)
        kti = kt.view(S, 6, 8, 16).transpose(1, 2); vti = vt.view(S, 6, 8, 16).transpose(1, 2)
        a = torch.softmax(qi @ kti.transpose(-1, -2) / 4.0, -1) @ vti
        cmp("a2", v["a2"], a.transpose(1, 2).reshape(S, hw, 128))
    if stop == 4:
        k1 = ln(keys + attn(lp + "cross_attn_image_to_token.", keys + kpe, q + tokens, q), lp + "norm4")
        cmp("keys after layer0", v["keysB"], k1)
