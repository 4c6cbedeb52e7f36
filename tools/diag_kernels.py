"""GPU bring-up diagnostics: runs each kernel on small and full shapes, prints error statistics.
Usage (on a B200): python tools/diag_kernels.py [gemm] [ln] [attn] [clip]"""
import os, sys, time, math, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import ops

torch.manual_seed(0)
dev = "cuda"

def stats(name, got, ref, tol):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-9
    mx = err.max().item()
    bad = (err > tol * denom)
    msg = f"{name}: max_abs_err={mx:.4e} ref_absmax={denom:.4e} rel={mx/denom:.3e} bad={int(bad.sum())}/{bad.numel()} nan={int(torch.isnan(got).sum())}"
    if bad.any():
        idx = bad.nonzero()
        rows = idx[:, 0].unique()[:12].tolist()
        cols = idx[:, 1].unique()[:12].tolist() if idx.shape[1] > 1 else []
        msg += f"\n    first bad rows {rows} cols {cols}; got={got[tuple(idx[0])].item():.4f} ref={ref[tuple(idx[0])].item():.4f}"
    print(("OK   " if not bad.any() and not torch.isnan(got).any() else "FAIL ") + msg, flush=True)
    return not bad.any()

def t_gemm():
    for (M, N, K) in [(128, 128, 64), (128, 256, 64), (256, 128, 128), (300, 384, 1024), (1025 * 2, 3072, 1024), (4096, 1024, 4096), (777, 512, 4096), (65600, 1024, 1024)]:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device=dev)
        ref = a.float() @ w.float().T + b
        for act, fn in [(ops.ACT_NONE, lambda x: x), (ops.ACT_QUICK_GELU, lambda x: x * torch.sigmoid(1.702 * x)), (ops.ACT_GELU_ERF, torch.nn.functional.gelu), (ops.ACT_RELU, torch.relu)]:
            try:
                out = ops.gemm(a, w, b, act=act, out_mode=ops.OUT_BF16)
                torch.cuda.synchronize()
                stats(f"gemm bf16 M{M} N{N} K{K} act{act}", out, fn(ref), 1e-2)
            except Exception as e:
                print("EXC", M, N, K, act, repr(e)); traceback.print_exc(); return
            if M > 5000: break
        res = torch.randn(M, N, device=dev)
        out = res.clone()
        ops.gemm(a, w, b, out_mode=ops.OUT_F32, out=out, resid=out)
        torch.cuda.synchronize()
        stats(f"gemm f32+resid M{M} N{N} K{K}", out, ref + res, 2e-3)
        out = ops.gemm(a, w, None, out_mode=ops.OUT_F32)
        torch.cuda.synchronize()
        stats(f"gemm f32 nobias M{M} N{N} K{K}", out, ref - b, 2e-3)
    # LN epilogue + periodic bias
    M, N, K = 3 * 1024, 256, 128
    a = (torch.randn(M, K, device=dev)).bfloat16(); w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev).bfloat16()
    g = torch.randn(N, device=dev); be = torch.randn(N, device=dev)
    out = ops.gemm(a, w, b, out_mode=ops.OUT_BF16_LN, resid=r, ln_gamma=g, ln_beta=be, ln_eps=1e-5)
    ref = torch.nn.functional.layer_norm(a.float() @ w.float().T + b + r.float(), (N,), g, be, 1e-5)
    torch.cuda.synchronize(); stats("gemm bf16+LN", out, ref, 1e-2)
    b2 = torch.randn(1024, 384, device=dev); w2 = (torch.randn(384, 256, device=dev) / 16).bfloat16(); a2 = torch.randn(M, 256, device=dev).bfloat16()
    out = ops.gemm(a2, w2, b2, bias_period=1024)
    ref = a2.float() @ w2.float().T + b2.repeat(3, 1)
    torch.cuda.synchronize(); stats("gemm periodic bias", out, ref, 1e-2)
    # timing of the big shapes
    for (M, N, K, act) in [(65600, 3072, 1024, 0), (65600, 4096, 1024, 1), (65600, 1024, 4096, 0), (65536, 8192, 1024, 2), (65536, 4096, 8192, 0)]:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16(); b = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(2): ops.gemm(a, w, b, act=act, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.gemm(a, w, b, act=act, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"time gemm M{M} N{N} K{K} act{act}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        for _ in range(2): torch.matmul(a, w.T)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5): torch.matmul(a, w.T)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"     cublas same shape: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)

def t_ln():
    for rows, D, dt in [(1000, 1024, torch.float32), (777, 256, torch.bfloat16), (65, 4096, torch.float32), (33, 5120, torch.bfloat16)]:
        x = (torch.randn(rows, D, device=dev) * 2 + 0.5).to(dt)
        g = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
        out = ops.layernorm(x, g, b, 1e-5)
        ref = torch.nn.functional.layer_norm(x.float(), (D,), g, b, 1e-5)
        torch.cuda.synchronize(); stats(f"layernorm {rows}x{D} {dt}", out, ref, 1e-2)

def attn_ref(qkv, heads, scale, kv=None):
    B, T, _ = qkv.shape
    q, k, v = qkv.float().view(B, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    if kv is not None:
        s = s.masked_fill(~kv.bool()[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, heads * 64)

def t_attn():
    for (B, T, H, masked) in [(1, 128, 1, False), (1, 256, 2, False), (2, 1025, 16, False), (2, 1025, 16, True), (1, 300, 3, True),
                              (2, 130, 2, True), (1, 5, 2, False), (1, 264, 1, False), (1, 1024, 4, False)]:
        qkv = torch.randn(B, T, 3 * H * 64, device=dev).bfloat16()
        kv = None
        if masked:
            kv = (torch.rand(B, T, device=dev) > 0.3).to(torch.uint8); kv[:, 0] = 1
        try:
            out = ops.attention_d64(qkv, H, 0.125, kv)
            torch.cuda.synchronize()
        except Exception as e:
            print("EXC attn", B, T, H, masked, repr(e)); return
        ref = attn_ref(qkv, H, 0.125, kv)
        stats(f"attention B{B} T{T} H{H} masked={masked}", out.reshape(B * T, -1), ref.reshape(B * T, -1), 2e-2)
    B, T, H = 64, 1025, 16
    qkv = torch.randn(B, T, 3 * H * 64, device=dev).bfloat16()
    out = torch.empty(B, T, H * 64, device=dev, dtype=torch.bfloat16)
    for _ in range(2): ops.attention_d64(qkv, H, 0.125, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.attention_d64(qkv, H, 0.125, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"time attention B64 T1025 H16: {ms:.3f} ms  {4*B*H*T*T*64/ms/1e9:.1f} TFLOP/s", flush=True)
    q, k, v = qkv.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    for _ in range(2): torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): torch.nn.functional.scaled_dot_product_attention(q, k, v)
    e1.record(); torch.cuda.synchronize()
    print(f"     torch sdpa: {e0.elapsed_time(e1)/5:.3f} ms", flush=True)

def t_clip():
    from walkgpt_b200.modules import CLIPVisionTower
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import path_a
    for layers, B in [(2, 2), (24, 2)]:
        m = CLIPVisionTower(layers=layers).cuda()
        px = torch.randn(B, 3, 448, 448, device=dev)
        sd = {k: v.float().cpu() for k, v in m.state_dict().items()}
        last, mid = m(px)
        torch.cuda.synchronize()
        n_states = layers + 1
        hs = path_a.clip_hidden_states(sd, px.cpu(), None, n_layers=max(m.hidden_state_indices()))
        il, im = m.hidden_state_indices()
        stats(f"clip {layers}L last(hs[{il}])", last.reshape(-1, 1024).cpu(), hs[il][:, 1:].reshape(-1, 1024), 2e-2)
        stats(f"clip {layers}L mid(hs[{im}])", mid[0].reshape(-1, 1024).cpu(), hs[im][:, 1:].reshape(-1, 1024), 2e-2)
    m = CLIPVisionTower(layers=24).cuda()
    px = torch.randn(64, 3, 448, 448, device=dev).bfloat16()
    for _ in range(2): m(px)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): m(px)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"time clip tower B64 (23 layers): {ms:.2f} ms -> {64/ms*1e3:.1f} img/s, {64*693.5/ms:.1f} TFLOP/s", flush=True)


def _sd_cpu(m):
    return {k: v.detach().float().cpu() for k, v in m.state_dict().items()}

def t_ctp():
    from walkgpt_b200.modules import CalibratedTextProjector
    from oracle import path_a
    for H, rows in [(4096, 192), (5120, 33), (256, 7)]:
        m = CalibratedTextProjector(H, 256, seed=3).cuda()
        x = torch.randn(rows, H, device=dev)
        y = m(x[None])[0]
        ref = path_a.ctp_forward(_sd_cpu(m), x.cpu()[None])[0]
        torch.cuda.synchronize(); stats(f"ctp H{H} rows{rows} fp32-in", y.cpu(), ref, 1e-2)
        yb = m(x.bfloat16()[None])[0]
        stats(f"ctp H{H} rows{rows} bf16-in", yb.float().cpu(), ref, 2e-2)
    y2 = m(x)
    print("ctp 2-D input shape", tuple(y2.shape))

def t_msqp():
    from walkgpt_b200.modules import MultiScaleQFormerProjector
    from oracle import path_a
    for sam_dim, H, B, L in [(256, 64, 2, 64), (1024, 4096, 2, 1024), (256, 128, 1, 4096)]:
        m = MultiScaleQFormerProjector(sam_dim, H, pad_to_square=True, target_square_side=6, seed=4).cuda()
        x = torch.randn(B, L, sam_dim, device=dev)
        y = m(x)
        torch.cuda.synchronize()
        ref = path_a.msqp_forward(_sd_cpu(m), x.bfloat16().float().cpu(), target_square_side=6)
        stats(f"msqp sam{sam_dim} H{H} B{B} L{L}", y.reshape(-1, H).cpu(), ref.reshape(-1, H), 2e-2)

def t_projneck():
    from walkgpt_b200.modules import ProjectorNeck
    from oracle import path_a
    for mm, H, B, g in [(128, 64, 2, 8), (1024, 4096, 2, 32)]:
        m = ProjectorNeck(mm, H, 256, seed=5).cuda()
        x = torch.randn(B, g * g, mm, device=dev)
        sd = _sd_cpu(m)
        sdp = {k[len("out_mm_projector."):]: v for k, v in sd.items() if k.startswith("out_mm_projector.")}
        sdn = {k[len("image_feature_neck."):]: v for k, v in sd.items() if k.startswith("image_feature_neck.")}
        p = m.project(x); e = m(x)
        torch.cuda.synchronize()
        pr = path_a.out_mm_projector_mlp(sdp, x.bfloat16().float().cpu())
        er = path_a.image_feature_neck(sdn, pr, g)
        stats(f"projector mm{mm} H{H}", p.reshape(-1, H).cpu(), pr.reshape(-1, H), 2e-2)
        stats(f"proj+neck mm{mm} H{H}", e.reshape(B * 256, -1).cpu(), er.reshape(B * 256, -1), 3e-2)
        e2 = m.neck(pr.permute(0, 2, 1).reshape(B, H, g, g).to(dev))
        stats(f"neck alone H{H}", e2.reshape(B * 256, -1).cpu(), er.reshape(B * 256, -1), 3e-2)

def t_decoder():
    from walkgpt_b200.modules import PromptEncoder, MaskDecoderMultiScale
    from oracle import path_a
    for g, S in [(8, 3), (32, 2), (32, 12)]:
        pe_m = PromptEncoder(256, (g, g), (g * 14, g * 14), 16, seed=6).cuda()
        dec = MaskDecoderMultiScale(seed=7).cuda()
        emb = torch.randn(1, 256, g, g, device=dev)
        txt = torch.randn(S, 1, 256, device=dev) * 0.5
        sparse, dense = pe_m(None, None, None, txt)
        pe = pe_m.get_dense_pe()
        sdp, sdd = _sd_cpu(pe_m), _sd_cpu(dec)
        pe_ref = path_a.dense_pe(sdp["pe_layer.positional_encoding_gaussian_matrix"], g, g)[None]
        stats(f"dense_pe g{g}", pe.reshape(256, -1).cpu(), pe_ref.reshape(256, -1), 1e-4)
        for mm in (False, True):
            masks, iou = dec(emb, pe, sparse, dense, mm, 0)
            torch.cuda.synchronize()
            rm, ri = path_a.mask_decoder_multiscale(sdd, emb.bfloat16().float().cpu(), pe_ref, sparse.cpu(), dense.cpu(), mm, 0)
            stats(f"decoder g{g} S{S} multimask={mm} masks", masks.reshape(masks.shape[0] * masks.shape[1], -1).cpu(), rm.reshape(rm.shape[0] * rm.shape[1], -1), 2e-2)
            stats(f"decoder g{g} S{S} multimask={mm} iou", iou.cpu(), ri, 2e-2)

def t_post():
    from walkgpt_b200.modules import postprocess_masks_fused, postprocess_masks
    from oracle import path_a
    low = torch.randn(5, 1, 64, 64, device=dev) * 3
    for inp, orig in [((448, 448), (448, 448)), ((252, 448), (360, 640)), ((448, 301), (517, 347))]:
        lg, mk, sc = postprocess_masks_fused(low[:, 0].contiguous(), inp, orig)
        torch.cuda.synchronize()
        ref = path_a.postprocess_masks(low.cpu(), inp, orig)[:, 0]
        stats(f"postprocess {inp}->{orig} logits", lg.reshape(5, -1).cpu(), ref.reshape(5, -1), 1e-5)
        stats(f"postprocess {inp}->{orig} score", sc.cpu()[None], path_a.mask_score(ref)[None], 1e-4)
        mism = (mk.cpu().bool() != (ref > 0)).sum().item()
        print(f"     mask mismatches: {mism} / {mk.numel()}")
    low_b = torch.randn(2, 1, 256, 256, device=dev) * 3
    lg, _, _ = postprocess_masks_fused(low_b[:, 0].contiguous(), (768, 1024), (480, 640), target_size=1024)
    ref = path_a.postprocess_masks(low_b.cpu(), (768, 1024), (480, 640), target_size=1024, cast_back=False)[:, 0]
    stats("postprocess SAM 256->1024->crop->480x640", lg.reshape(2, -1).cpu(), ref.reshape(2, -1), 1e-5)
    low = torch.randn(64 * 12, 64, 64, device=dev)
    for _ in range(3): postprocess_masks_fused(low, (448, 448), (448, 448))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): postprocess_masks_fused(low, (448, 448), (448, 448))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = 768 * (64 * 64 * 4 + 448 * 448 * 5)
    print(f"time postprocess 768 masks (incl. torch.empty): {ms:.3f} ms -> {byts/ms/1e6:.1f} GB/s", flush=True)
    low = torch.randn(64 * 3, 64, 64, device=dev)
    for _ in range(3): postprocess_masks_fused(low, (448, 448), (448, 448))
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): postprocess_masks_fused(low, (448, 448), (448, 448))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = 192 * (64 * 64 * 4 + 448 * 448 * 5)
    print(f"time postprocess 192 masks (incl. torch.empty): {ms:.3f} ms -> {byts/ms/1e6:.1f} GB/s", flush=True)

def t_path():
    from walkgpt_b200.modules import GroundingPath, merge_split
    from oracle import path_a
    B, S, H = 2, 3, 4096
    m = GroundingPath(hidden_size=H, clip_layers=24, seed=1).cuda()
    px = torch.randn(B, 3, 448, 448, device=dev)
    seg = torch.randn(B * S, H, device=dev)
    offs = [0, S, 2 * S]
    out = m(px, seg, offs)
    torch.cuda.synchronize()
    def sub(prefix):
        mod = dict(m.named_children())
        return None
    W = {
        "clip": _sd_cpu(m.vision_tower), "msqp": _sd_cpu(m.msqp),
        "proj": {k[len("out_mm_projector."):]: v for k, v in _sd_cpu(m.proj_neck).items() if k.startswith("out_mm_projector.")},
        "neck": {k[len("image_feature_neck."):]: v for k, v in _sd_cpu(m.proj_neck).items() if k.startswith("image_feature_neck.")},
        "ctp": _sd_cpu(m.text_hidden_fcs[0]), "prompt": _sd_cpu(m.prompt_encoder), "decoder": _sd_cpu(m.mask_decoder),
    }
    ref = path_a.path_a_forward(W, px.bfloat16().float().cpu(), seg.cpu(), offs)
    stats("path vis_tokens", out["vis_tokens"].reshape(-1, H).float().cpu(), ref["vis_tokens"].reshape(-1, H), 3e-2)
    stats("path txt_emb", out["txt_emb"].cpu(), ref["txt_emb"], 2e-2)
    emb_ref = ref["img_emb"].flatten(2).permute(0, 2, 1).reshape(-1, 256)
    stats("path img_emb", merge_split(out["img_emb_split"]).reshape(-1, 256).cpu(), emb_ref, 3e-2)
    stats("path low_res", out["low_res"].reshape(B * S, -1).cpu(), ref["low_res"].reshape(B * S, -1), 3e-2)
    stats("path iou", out["iou"].cpu(), ref["iou"], 3e-2)
    stats("path logits", out["logits"].reshape(B * S, -1).cpu(), ref["logits"].reshape(B * S, -1), 3e-2)
    err = (out["logits"].cpu() - ref["logits"]).abs().max().item()
    a, b = out["masks"].cpu().bool(), ref["logits"] > 0
    iou = ((a & b).flatten(1).sum(1).float() / ((a | b).flatten(1).sum(1).float() + 1e-9))
    print(f"     logits max-abs err {err:.4e} (gate 2e-2); logit absmax {ref['logits'].abs().max().item():.3f}; per-mask IoU {iou.tolist()} (gate 0.995)")
    print("     scores", out["scores"].cpu().tolist(), ref["scores"].tolist())
    print("     depth", out["depth"].cpu().tolist())
    # throughput, config 2: B=64, S=3
    B, S = 64, 3
    px = torch.randn(B, 3, 448, 448, device=dev).bfloat16(); seg = torch.randn(B * S, H, device=dev).bfloat16()
    offs = list(range(0, B * S + 1, S))
    for _ in range(2): m(px, seg, offs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): m(px, seg, offs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"time full path B64 S3: {ms:.2f} ms -> {B/ms*1e3:.1f} img/s ({B*800.7/ms:.1f} TFLOP/s)", flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "ln", "attn", "clip"]
    print(torch.cuda.get_device_name(0), flush=True)
    for w in which:
        print(f"===== {w} =====", flush=True)
        try:
            {"gemm": t_gemm, "ln": t_ln, "attn": t_attn, "clip": t_clip, "ctp": t_ctp, "msqp": t_msqp, "projneck": t_projneck, "decoder": t_decoder, "post": t_post, "path": t_path}[w]()
        except Exception as e:
            print("EXC in", w, repr(e)); traceback.print_exc()
