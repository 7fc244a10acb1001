"""Pre-training twin of the planner on B200.

``B200TextPathCMT`` mirrors ``GlocalTextPathCMT`` (pretrain_src/pretrain_src/model/vilmodel.py:656-754): the same
layers as the navigation model driven in TRAJECTORY form — ``ImageEmbeddings.forward`` over the (sum of steps) x views
batch (:488-534), ``GlobalMapEncoder._aggregate_gmap_features`` (:585-619), the cross-modal encoder in both directions
(``GraphLXRTXLayer.forward`` :384-398 and ``forward_lang2visn`` :400-411).  ``B200PreTraining`` mirrors
``GlocalTextPathCMTPreTraining`` (pretrain_src/pretrain_src/model/pretrain_cmt.py:50-283) for the two tasks the
reference pre-trains R2R-CE with (run_pt/r2r_pretrain_habitat.json: ``mlm``, ``sap``): same ``forward(batch, task,
compute_loss)`` entry, same ``state_dict()`` keys (``bert.*``, ``mlm_head.*`` with the tied decoder, ``global_sap_head.*``).

Everything numeric runs in the sm_100a kernels of ``libetpnav_b200.so`` through the C ABI: the trajectory form reuses
``etp_forward_txt`` / ``etp_forward_panorama`` / ``etp_forward_navigation`` (one call each per batch), the aggregation
and the masked-token gather are ``etp_segment_gather`` over a host-built CSR (the reference's string-keyed dictionaries
become integer index lists; no tensor arithmetic happens on the host), ``forward_mlm`` is ``etp_forward_lang2visn``, the
MLM head is two ``etp_gemm`` + ``etp_layernorm_fwd``.  No CPU / PyTorch fallback exists.
Not built: object features (``obj_feat_size`` is 0 in the reference's R2R-CE config), the ``mrc`` / ``og`` tasks (their
code path in the reference calls ``self.bert`` with a stale signature and is not used by ``r2r_pretrain_habitat.json``).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib as _L
from .config import PlannerConfig
from .planner import (B200Planner, NavWeights, LayerWeights, _f32c, _nav_inputs, _param_grads, i32, p_void)

_declared = False


def _declare():
    global _declared
    if _declared:
        return
    L = _L.lib()
    L.etp_segment_gather.argtypes = [p_void, p_void, p_void, p_void, i32, i32, p_void, p_void]
    L.etp_l2v_saved_bytes.restype = C.c_size_t
    L.etp_l2v_saved_bytes.argtypes = [i32] * 5
    L.etp_l2v_bwd_work_bytes.restype = C.c_size_t
    L.etp_l2v_bwd_work_bytes.argtypes = [i32] * 4
    L.etp_forward_lang2visn.argtypes = [C.POINTER(NavWeights), C.c_void_p, p_void, p_void, C.c_size_t, i32, p_void]
    L.etp_backward_lang2visn.argtypes = [C.POINTER(NavWeights), C.POINTER(NavWeights), C.c_void_p, p_void, p_void,
                                         C.c_size_t, p_void, C.c_size_t, p_void, p_void, p_void]
    _declared = True


# ----------------------------------------------------------------------------------------------------
# host side of _aggregate_gmap_features: dictionaries of strings -> CSR index lists (pure bookkeeping)
# ----------------------------------------------------------------------------------------------------
class Csr:
    """seg_ptr [S+1] int32, index [nnz] int32, weight [nnz] fp32 (CPU numpy until ``.to(device)``)."""

    def __init__(self, seg_ptr, index, weight, num_src):
        self.seg_ptr, self.index, self.weight, self.num_src = seg_ptr, index, weight, num_src
        self.num_segments = len(seg_ptr) - 1

    def transposed(self):
        """CSR of the adjoint map (source row -> output rows that read it): the backward is the same gather."""
        seg = np.repeat(np.arange(self.num_segments, dtype=np.int32), np.diff(self.seg_ptr))
        order = np.argsort(self.index, kind="stable")
        counts = np.bincount(self.index, minlength=self.num_src)
        ptr = np.zeros(self.num_src + 1, dtype=np.int32)
        np.cumsum(counts, out=ptr[1:])
        return Csr(ptr, seg[order].astype(np.int32), self.weight[order].astype(np.float32), self.num_segments)

    def to(self, device):
        d = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt, non_blocking=True)
        return d(self.seg_ptr, torch.int32), d(self.index, torch.int32), d(self.weight, torch.float32)


def build_gmap_csr(traj_step_lens, traj_vp_view_lens, traj_vpids, traj_cand_vpids, gmap_vpids, view_pitch, n_max=None):
    """Index form of ``GlobalMapEncoder._aggregate_gmap_features`` (vilmodel.py:585-619).

    Source rows are the view tokens of the trajectory batch, row = (global step) * view_pitch + view; output rows are
    b * n_max + n over the padded map.  Row n = 0 ([stop], :613-617) and the padding rows stay empty (zeros).  A visited
    viewpoint is the mean over its valid views at its LAST visit (the dictionary entry is overwritten, :599); an
    unvisited one the mean of every candidate token that pointed at it before it was visited (:600-603)."""
    lens = [int(x) for x in traj_vp_view_lens]
    B = len(gmap_vpids)
    if n_max is None:
        n_max = max(len(g) for g in gmap_vpids)
    ptr, idx, wt = [0], [], []
    step0 = 0
    for i in range(B):
        visited, unvisited = {}, {}
        for t in range(traj_step_lens[i]):
            visited[traj_vpids[i][t]] = step0 + t
            for j, vp in enumerate(traj_cand_vpids[i][t]):
                if vp not in visited:
                    unvisited.setdefault(vp, []).append((step0 + t) * view_pitch + j)
        for n in range(n_max):
            if 1 <= n < len(gmap_vpids[i]):
                vp = gmap_vpids[i][n]
                if vp in visited:
                    g = visited[vp]
                    k = lens[g]
                    idx.extend(range(g * view_pitch, g * view_pitch + k))
                    wt.extend([1.0 / k] * k)
                else:
                    src = unvisited[vp]
                    idx.extend(src)
                    wt.extend([1.0 / len(src)] * len(src))
            ptr.append(len(idx))
        step0 += traj_step_lens[i]
    return Csr(np.asarray(ptr, dtype=np.int32), np.asarray(idx, dtype=np.int32), np.asarray(wt, dtype=np.float32),
               step0 * view_pitch), n_max


def _segment_gather(src, ptr, idx, wt, num_segments):
    _declare()
    out = torch.empty(num_segments, src.shape[1], device=src.device, dtype=torch.float32)
    _L._check(_L.lib().etp_segment_gather(_L.ptr(src), _L.ptr(ptr), _L.ptr(idx), _L.ptr(wt), num_segments, src.shape[1],
                                          _L.ptr(out), _L.stream_ptr()), "etp_segment_gather")
    return out


class _SegGatherFn(torch.autograd.Function):
    """out = A . src for a sparse row map A given as CSR; backward = A^T . dout (the transposed CSR, same kernel)."""

    @staticmethod
    def forward(ctx, src, fwd, bwd, num_segments, num_src):
        ctx.bwd, ctx.num_src = bwd, num_src
        return _segment_gather(_f32c(src), *fwd, num_segments)

    @staticmethod
    def backward(ctx, dout):
        return _segment_gather(_f32c(dout), *ctx.bwd, ctx.num_src), None, None, None, None


def segment_gather(src2d, csr: Csr):
    """Differentiable ``etp_segment_gather``: src2d fp32 [R, W] -> [csr.num_segments, W]."""
    assert src2d.shape[0] == csr.num_src
    dev = src2d.device
    return _SegGatherFn.apply(src2d, csr.to(dev), csr.transposed().to(dev), csr.num_segments, csr.num_src)


# ----------------------------------------------------------------------------------------------------
# lang2visn step call
# ----------------------------------------------------------------------------------------------------
def _l2v_forward(m, txt, img, aux, training, drop=None):
    _declare()
    B, N = img.shape[:2]
    Lt = txt.shape[1]
    L = _L.lib()
    ni = _nav_inputs(txt, img, aux, drop)
    out = torch.empty(B, Lt, 768, device=img.device, dtype=torch.float32)
    nbytes = L.etp_l2v_saved_bytes(B, N, Lt, m.config.num_x_layers, training)
    saved = torch.empty(nbytes, dtype=torch.uint8, device=img.device)
    _L._check(L.etp_forward_lang2visn(C.byref(m._structs["l2v"]), C.byref(ni), _L.ptr(out), _L.ptr(saved), nbytes,
                                      training, _L.stream_ptr()), "etp_forward_lang2visn")
    return out, saved


class _L2VFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m, txt_embeds, gmap_img_fts, aux, drop, anchor, *params):
        txt, img = _f32c(txt_embeds), _f32c(gmap_img_fts)
        out, saved = _l2v_forward(m, txt, img, aux, 1, drop)
        ctx.m, ctx.saved, ctx.keep, ctx.nparams, ctx.drop = m, saved, (txt, img, aux), len(params), drop
        return out

    @staticmethod
    def backward(ctx, d_out):
        m = ctx.m
        txt, img, aux = ctx.keep
        L = _L.lib()
        B, N, Lt = img.shape[0], img.shape[1], txt.shape[1]
        gbuf, gstart, per_call = m._l2v_grad_target()
        gst = m._grad_structs_for(gbuf, gstart)
        ni = _nav_inputs(txt, img, aux, ctx.drop)
        d_txt = torch.empty_like(txt) if ctx.needs_input_grad[1] else None
        d_img = torch.empty_like(img) if ctx.needs_input_grad[2] else None
        wbytes = L.etp_l2v_bwd_work_bytes(B, N, Lt, m.config.num_x_layers)
        work = torch.empty(wbytes, dtype=torch.uint8, device=img.device)
        _L._check(L.etp_backward_lang2visn(C.byref(m._structs["l2v"]), C.byref(gst["l2v"]), C.byref(ni),
                                           _L.ptr(_f32c(d_out)), _L.ptr(ctx.saved), ctx.saved.numel(), _L.ptr(work), wbytes,
                                           _L.ptr(d_txt), _L.ptr(d_img), _L.stream_ptr()), "etp_backward_lang2visn")
        pg = _param_grads(m, m._l2v_param_names(), gbuf, gstart, per_call, ctx.nparams)
        return (None, d_txt, d_img, None, None, None, *pg)


# ----------------------------------------------------------------------------------------------------
# the twin
# ----------------------------------------------------------------------------------------------------
def _seq_masks(lens, width, device):
    """gen_seq_masks (pretrain_src/pretrain_src/model/ops.py:37-45) with the padded width known from the data tensor
    (the reference takes max(lens), which the collate functions make equal to it): no host sync."""
    return torch.arange(width, device=device)[None] < lens.to(device)[:, None]


class B200TextPathCMT(B200Planner):
    """Replacement for ``GlocalTextPathCMT`` (pretrain_src/pretrain_src/model/vilmodel.py:656)."""

    def __init__(self, config: PlannerConfig, device="cuda"):
        super().__init__(config, device=device)

    # ------------------------------------------------------------------ structs / gradient plumbing
    def _build_structs(self, b32, b16, s16):
        s = super()._build_structs(b32, b16, s16)
        cfg = self.config
        if cfg.use_lang2visn_attn and cfg.num_x_layers > 0:
            X = cfg.num_x_layers
            ll = (LayerWeights * X)()
            for i in range(X):
                p = f"global_encoder.encoder.x_layers.{i}."
                ll[i] = self._layer_struct(p + "lang_self_att.self.", p + "lang_self_att.output.", p + "lang_inter.",
                                           p + "lang_output.", cross=p + "visual_attention.")
            nw, src = NavWeights(), s["nav"]
            nw.num_x_layers, nw.ln_eps, nw.layers = X, cfg.layer_norm_eps, ll
            for f in ("pos_w", "pos_b", "pos_g", "pos_bb", "step_emb", "xkv_all_w", "xkv_all_b"):
                setattr(nw, f, getattr(src, f))
            s["l2v"], s["l2v_layers"] = nw, ll
        return s

    def _l2v_param_names(self):
        """Parameters forward_mlm's cross-modal part reads: the navigation group (node packing, visual_attention) and the
        lang_* blocks of the ``pre`` group (adjacent in the flat layout)."""
        return self._group_names("nav") + [n for n in self._group_names("pre") if ".lang_" in n]

    def _l2v_grad_target(self):
        if self._direct_grad is not None:
            return self._direct_grad, 0, False
        gs = self.layout.group_ranges["nav"][0]
        ge = self.layout.group_ranges["pre"][1]
        return torch.zeros(ge - gs, dtype=torch.float32, device=self._flat.device), gs, True

    # ------------------------------------------------------------------ pieces of the reference forward
    def traj_img_embeddings(self, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts, traj_nav_types,
                       traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens):
        """``ImageEmbeddings.forward`` (vilmodel.py:488-534) -> (split_traj_embeds, split_traj_vp_lens).
        The trajectory batch [sum(steps), V, *] is one ``etp_forward_panorama`` call."""
        if traj_obj_img_fts is not None:
            raise NotImplementedError("object features are not part of the R2R-CE pre-training configuration")
        embeds, _ = self.forward_panorama(traj_view_img_fts, traj_view_dep_fts, traj_loc_fts, traj_nav_types,
                                          traj_vp_view_lens.to(traj_view_img_fts.device))
        return torch.split(embeds, traj_step_lens, 0), torch.split(traj_vp_view_lens, traj_step_lens, 0)

    def aggregate_gmap_features(self, split_traj_embeds, split_traj_vp_lens, traj_vpids, traj_cand_vpids, gmap_vpids,
                                n_max=None):
        """``GlobalMapEncoder._aggregate_gmap_features`` (vilmodel.py:585-619) -> gmap_img_fts [B, Nmax, 768]."""
        traj = torch.cat(list(split_traj_embeds), 0) if not isinstance(split_traj_embeds, torch.Tensor) else split_traj_embeds
        V = traj.shape[1]
        lens = torch.cat([x.reshape(-1) for x in split_traj_vp_lens]).tolist()
        steps = [len(x) for x in split_traj_vp_lens]
        csr, n_max = build_gmap_csr(steps, lens, traj_vpids, traj_cand_vpids, gmap_vpids, V, n_max)
        out = segment_gather(traj.reshape(-1, traj.shape[2]), csr)
        return out.view(len(gmap_vpids), n_max, traj.shape[2])

    def _gmap_aux(self, txt_masks, gmap_step_ids, gmap_pos_fts, gmap_lens, gmap_visited_masks, gmap_pair_dists):
        dev = self._flat.device
        N = gmap_step_ids.shape[1]
        gm = _seq_masks(gmap_lens, N, dev)
        if gmap_visited_masks is None:
            gmap_visited_masks = torch.zeros(gmap_step_ids.shape, dtype=torch.bool, device=dev)
        if gmap_pair_dists is None:
            gmap_pair_dists = torch.zeros(gmap_step_ids.shape[0], N, N, device=dev)
        return (self._mask_u8(txt_masks), gmap_step_ids.to(dev).contiguous().long(), _f32c(gmap_pos_fts.to(dev)),
                self._mask_u8(gm), self._mask_u8(gmap_visited_masks.to(dev)), _f32c(gmap_pair_dists.to(dev))), gm

    def _encode(self, txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts,
                traj_nav_types, traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids,
                gmap_vpids, n_max):
        """Text + trajectory embedding + map aggregation shared by ``forward`` and ``forward_mlm`` (vilmodel.py:673-686)."""
        txt_masks = _seq_masks(txt_lens, txt_ids.shape[1], txt_ids.device)
        txt_embeds = self.forward_txt(txt_ids, txt_masks)  # token_type_ids = 0 (:673-676)
        split_embeds, split_lens = self.traj_img_embeddings(traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts,
                                                       traj_nav_types, traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens)
        gmap_img_fts = self.aggregate_gmap_features(split_embeds, split_lens, traj_vpids, traj_cand_vpids, gmap_vpids, n_max)
        return txt_embeds, txt_masks, gmap_img_fts

    def forward_gmap(self, txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts,
                     traj_nav_types, traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids,
                     gmap_lens, gmap_step_ids, gmap_pos_fts, gmap_pair_dists, gmap_vpids, gmap_visited_masks=None):
        """The twin's ``forward`` plus the SAP logits of ``forward_sap`` (pretrain_cmt.py:218-262): dict with
        ``gmap_embeds`` [B,N,768] and ``global_logits`` [B,N] (-inf at visited / padded nodes)."""
        txt_embeds, txt_masks, gmap_img_fts = self._encode(
            txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts, traj_nav_types,
            traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids, gmap_vpids,
            gmap_step_ids.shape[1])
        dev = self._flat.device
        gm = _seq_masks(gmap_lens, gmap_step_ids.shape[1], dev)
        if gmap_visited_masks is None:
            gmap_visited_masks = torch.zeros(gmap_step_ids.shape, dtype=torch.bool, device=dev)
        if not self.config.graph_sprels or gmap_pair_dists is None:
            gmap_pair_dists = torch.zeros(gmap_step_ids.shape[0], gmap_step_ids.shape[1], gmap_step_ids.shape[1], device=dev)
        return self.forward_navigation(txt_embeds, txt_masks, gmap_vpids, gmap_step_ids.to(dev), gmap_img_fts,
                                       gmap_pos_fts.to(dev), gm, gmap_visited_masks.to(dev), gmap_pair_dists.to(dev))

    def forward(self, txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts,
                traj_nav_types, traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids,
                gmap_lens, gmap_step_ids, gmap_pos_fts, gmap_pair_dists, gmap_vpids, return_gmap_embeds=True):
        """``GlocalTextPathCMT.forward`` (vilmodel.py:668-711) -> gmap_embeds [B,N,768]."""
        if not return_gmap_embeds:
            return None
        return self.forward_gmap(txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts,
                                 traj_nav_types, traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids,
                                 traj_cand_vpids, gmap_lens, gmap_step_ids, gmap_pos_fts, gmap_pair_dists,
                                 gmap_vpids)["gmap_embeds"]

    def forward_mlm(self, txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts,
                    traj_nav_types, traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids,
                    gmap_lens, gmap_step_ids, gmap_pos_fts, gmap_pair_dists, gmap_vpids):
        """``GlocalTextPathCMT.forward_mlm`` (vilmodel.py:713-754) -> txt_embeds [B,L,768] after the lang2visn layers."""
        if not self.config.use_lang2visn_attn:
            raise ValueError("forward_mlm needs use_lang2visn_attn (the lang_* blocks of the x-layers)")
        txt_embeds, txt_masks, gmap_img_fts = self._encode(
            txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts, traj_nav_types,
            traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids, gmap_vpids,
            gmap_step_ids.shape[1])
        aux, _ = self._gmap_aux(txt_masks, gmap_step_ids, gmap_pos_fts, gmap_lens, None, None)
        self._refresh_cache()
        drop = self._next_dropout()
        if torch.is_grad_enabled() and (txt_embeds.requires_grad or gmap_img_fts.requires_grad
                                        or any(self._pmap[n].requires_grad for n in self._l2v_param_names())):
            params = [] if self._direct_grad is not None else [self._pmap[n] for n in self._l2v_param_names()]
            return _L2VFn.apply(self, txt_embeds, gmap_img_fts, aux, drop, self._anchor_t(), *params)
        return _l2v_forward(self, _f32c(txt_embeds), _f32c(gmap_img_fts), aux, 0, drop)[0]


# ----------------------------------------------------------------------------------------------------
# MLM head
# ----------------------------------------------------------------------------------------------------
def _pad64(n):
    return (n + 63) // 64 * 64


class _MlmHeadFn(torch.autograd.Function):
    """``BertOnlyMLMHead`` (vilmodel.py:258-299): dense -> gelu -> LayerNorm -> decoder (tied to the word embeddings,
    pretrain_cmt.py:79-82) + bias.  Rows = the masked tokens only (pretrain_cmt.py:148-149).  Two tcgen05 GEMMs with
    fused bias / GELU epilogues and one LayerNorm kernel forward; dgrad / wgrad GEMMs backward."""

    @staticmethod
    def forward(ctx, m, hidden, dense_w, dense_b, ln_g, ln_b, dec_bias, word_emb):
        x = _f32c(hidden)
        M, H = x.shape
        V = m.config.vocab_size
        dev = x.device
        w16 = lambda n: m._flat_bf16[m.layout.offset(n): m.layout.offset(n) + m.layout.entries[n][1]].view(m.layout.entries[n][2])
        Wd, Wemb = w16("mlm_head.predictions.transform.dense.weight"), w16("embeddings.word_embeddings.weight")
        xb = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
        _L.cast_bf16(x, xb)
        g = torch.empty(M, H, device=dev, dtype=torch.float32)
        gp = torch.empty(M, H, device=dev, dtype=torch.bfloat16)        # gelu'(pre-activation), for the backward
        _L.gemm(xb, Wd, bias=dense_b.detach(), act=1, out_f32=g, out_pre=gp, pre_mode=1)
        hb = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
        mean = torch.empty(M, device=dev, dtype=torch.float32)
        rstd = torch.empty(M, device=dev, dtype=torch.float32)
        _L.layernorm_fwd(g, ln_g.detach(), ln_b.detach(), m.config.layer_norm_eps, y_bf16=hb, mean=mean, rstd=rstd)
        ldp = _pad64(V)
        logits = torch.empty(M, ldp, device=dev, dtype=torch.float32)   # row pitch padded to 64 floats
        _L.gemm(hb, Wemb, bias=dec_bias.detach(), out_f32=logits, N=V)
        ctx.m, ctx.keep = m, (xb, g, gp, hb, mean, rstd, Wd, Wemb, ln_g.detach())
        return logits[:, :V]

    @staticmethod
    def backward(ctx, d_logits):
        m = ctx.m
        xb, g, gp, hb, mean, rstd, Wd, Wemb, ln_g = ctx.keep
        M, H = xb.shape
        V = m.config.vocab_size
        dev = xb.device
        ldp = _pad64(V)
        direct = m._direct_grad is not None   # PretrainTrainer: accumulate straight into the flat gradient buffer

        def target(name, shape):
            if direct:
                off, numel, shp = m.layout.entries[name]
                return m._direct_grad[off:off + numel].view(shp)
            return torch.zeros(shape, device=dev, dtype=torch.float32)

        dl = d_logits.contiguous().float()
        dlb = torch.zeros(M, ldp, device=dev, dtype=torch.bfloat16)
        dlb[:, :V].copy_(dl)
        d_bias = target("mlm_head.predictions.bias", (V,))
        _L.colsum(dl, d_bias)                                                               # += column sums
        d_emb = target("embeddings.word_embeddings.weight", (V, H))
        _L.gemm(dlb, hb, a_mn=True, b_mn=True, out_f32=d_emb, resid=d_emb, M=V, N=H, K=M)  # dW_dec += dlogits^T . h
        dh = torch.empty(M, H, device=dev, dtype=torch.float32)
        _L.gemm(dlb, Wemb, b_mn=True, out_f32=dh, M=M, N=H, K=V)                           # dh = dlogits . W_dec
        dg = torch.empty(M, H, device=dev, dtype=torch.float32)
        d_gamma = target("mlm_head.predictions.transform.LayerNorm.weight", (H,))
        d_beta = target("mlm_head.predictions.transform.LayerNorm.bias", (H,))
        _L.layernorm_bwd(dh, g, ln_g, mean, rstd, dg, dgamma=d_gamma, dbeta=d_beta)
        dtb = (dg * gp.float()).to(torch.bfloat16)                                           # * gelu'(pre): [M,768] glue
        d_db = target("mlm_head.predictions.transform.dense.bias", (H,))
        _L.colsum(dtb, d_db)
        d_dw = target("mlm_head.predictions.transform.dense.weight", (H, H))
        _L.gemm(dtb, xb, a_mn=True, b_mn=True, out_f32=d_dw, resid=d_dw, M=H, N=H, K=M)
        dx = torch.empty(M, H, device=dev, dtype=torch.float32)
        _L.gemm(dtb, Wd, b_mn=True, out_f32=dx, M=M, N=H, K=H)
        if direct:
            return None, dx, None, None, None, None, None, None
        return None, dx, d_dw, d_db, d_gamma, d_beta, d_bias, d_emb


class B200PreTraining(nn.Module):
    """Replacement for ``GlocalTextPathCMTPreTraining`` (pretrain_src/pretrain_src/model/pretrain_cmt.py:50) for the
    ``mlm`` and ``sap`` tasks.  ``state_dict()`` uses the reference's keys: ``bert.<twin key>``, ``mlm_head.*`` (with
    ``mlm_head.predictions.decoder.weight`` as the tied alias of ``bert.embeddings.word_embeddings.weight``) and
    ``global_sap_head.*`` at the top level."""

    def __init__(self, config: PlannerConfig, device="cuda"):
        super().__init__()
        config.use_lang2visn_attn = True
        config.mlm_head = True
        self.config = config
        self.bert = B200TextPathCMT(config, device=device)
        self._register_state_dict_hook(self._sd_post_hook)
        self._register_load_state_dict_pre_hook(self._sd_load_pre_hook)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, config=None, state_dict=None, device="cuda"):
        """The constructor call of the reference's driver (pretrain_src/pretrain_src/train_r2r.py:146-148:
        ``model_class.from_pretrained(pretrained_model_name_or_path=None, config=model_config, state_dict=checkpoint)``).
        ``config`` is the reference's model config object (attributes of run_pt/r2r_model_config_dep.json) or a
        ``PlannerConfig``; keys of ``state_dict`` this model does not own are ignored, as HF does."""
        if pretrained_model_name_or_path is not None:
            raise ValueError("only the reference's own usage (name None + explicit state_dict) is supported")
        if not isinstance(config, PlannerConfig):
            fields = PlannerConfig.__dataclass_fields__
            kw = {k: getattr(config, k) for k in fields if hasattr(config, k)}
            if getattr(config, "depth_feat_size", 128) == 0:
                kw["use_depth_embedding"], kw["depth_feat_size"] = False, 128
            config = PlannerConfig(**kw)
        model = cls(config, device=device)
        if state_dict is not None:
            model.load_state_dict(state_dict, strict=False)
        return model

    # ------------------------------------------------------------------ reference key layout
    @staticmethod
    def _ref_key(k):
        return k if k.startswith(("mlm_head.", "global_sap_head.")) else "bert." + k

    # The reference's key layout is produced / consumed by nn.Module's own state_dict machinery through two hooks, so it
    # also holds when this model is a CHILD (DDP / DataParallel / any wrapper: train_r2r.py's ModelSaver calls
    # ``model.state_dict()`` on the wrapper and strips ``module.``, utils/save.py:23-46): a parent's recursion reaches the
    # hooks with its own ``destination`` / ``prefix``.
    @staticmethod
    def _sd_post_hook(module, state_dict, prefix, local_metadata):
        for head in ("mlm_head.", "global_sap_head."):
            src = prefix + "bert." + head
            for k in [k for k in state_dict if k.startswith(src)]:
                state_dict[prefix + head + k[len(src):]] = state_dict.pop(k)
        w = prefix + "bert.embeddings.word_embeddings.weight"
        if w in state_dict:   # tied decoder (pretrain_cmt.py:79-82)
            state_dict[prefix + "mlm_head.predictions.decoder.weight"] = state_dict[w]
        return state_dict

    @staticmethod
    def _sd_load_pre_hook(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        state_dict.pop(prefix + "mlm_head.predictions.decoder.weight", None)
        for k in [k for k in state_dict if k.startswith(prefix) and k.endswith("position_ids")]:
            state_dict.pop(k)
        for head in ("mlm_head.", "global_sap_head."):
            src = prefix + head
            for k in [k for k in state_dict if k.startswith(src)]:
                state_dict[prefix + "bert." + head + k[len(src):]] = state_dict.pop(k)

    def load_state_dict(self, sd, strict=True, **kw):
        """Keys this model does not own are ignored unless ``strict`` (HF ``from_pretrained(state_dict=...)`` semantics)."""
        own = set(self.state_dict().keys())
        sd = {k: v for k, v in sd.items() if not k.endswith("position_ids")}
        foreign = [k for k in sd if k not in own]
        if strict and foreign:
            raise KeyError(f"unexpected keys: {foreign[:5]}")
        out = super().load_state_dict({k: v for k, v in sd.items() if k in own}, strict=strict, **kw)
        self.bert._cache_key = None
        return out

    def set_dropout(self, p: float):
        """Replacement for ``set_dropout(model, opts.dropout)`` (train_r2r.py:150, utils/misc.py:19-25)."""
        self.bert.set_dropout(p)

    # ------------------------------------------------------------------ tasks
    def forward(self, batch, task, compute_loss=True):
        """pretrain_cmt.py:84-135."""
        b = batch
        common = (b["txt_ids"], b["txt_lens"], b["traj_view_img_fts"], b.get("traj_view_dep_fts"), b.get("traj_obj_img_fts"),
                  b["traj_loc_fts"], b["traj_nav_types"], b["traj_step_lens"], b["traj_vp_view_lens"],
                  b.get("traj_vp_obj_lens"), b["traj_vpids"], b["traj_cand_vpids"], b["gmap_lens"], b["gmap_step_ids"],
                  b["gmap_pos_fts"], b["gmap_pair_dists"], b["gmap_vpids"])
        if task.startswith("mlm"):
            return self.forward_mlm(*common, b["txt_labels"], compute_loss)
        if task.startswith("sap"):
            return self.forward_sap(*common, b["gmap_visited_masks"], b["global_act_labels"], b.get("local_act_labels"),
                                    compute_loss)
        raise ValueError("invalid task")

    def forward_mlm(self, txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts,
                    traj_nav_types, traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids,
                    gmap_lens, gmap_step_ids, gmap_pos_fts, gmap_pair_dists, gmap_vpids, txt_labels, compute_loss):
        """pretrain_cmt.py:137-158: MLM scores (or per-token CE) on the masked positions only."""
        m = self.bert
        txt_embeds = m.forward_mlm(txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts,
                                   traj_nav_types, traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids,
                                   traj_cand_vpids, gmap_lens, gmap_step_ids, gmap_pos_fts, gmap_pair_dists, gmap_vpids)
        B, Lt, H = txt_embeds.shape
        # _compute_masked_hidden (:160-164): rows of the masked tokens, in row-major order — a unit-weight segment gather
        pos = (txt_labels.reshape(-1) != -1).nonzero().reshape(-1).cpu().numpy().astype(np.int32)
        M = len(pos)
        if M == 0:
            return txt_embeds.new_zeros(0) if compute_loss else txt_embeds.new_zeros(0, m.config.vocab_size)
        csr = Csr(np.arange(M + 1, dtype=np.int32), pos, np.ones(M, dtype=np.float32), B * Lt)
        masked = segment_gather(txt_embeds.reshape(B * Lt, H), csr)
        pm = m._pmap
        scores = _MlmHeadFn.apply(m, masked, pm["mlm_head.predictions.transform.dense.weight"],
                                  pm["mlm_head.predictions.transform.dense.bias"],
                                  pm["mlm_head.predictions.transform.LayerNorm.weight"],
                                  pm["mlm_head.predictions.transform.LayerNorm.bias"], pm["mlm_head.predictions.bias"],
                                  pm["embeddings.word_embeddings.weight"])
        if compute_loss:
            labels = txt_labels.to(scores.device)
            return F.cross_entropy(scores, labels[labels != -1], reduction="none")
        return scores

    def forward_sap(self, txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts, traj_loc_fts,
                    traj_nav_types, traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids,
                    gmap_lens, gmap_step_ids, gmap_pos_fts, gmap_pair_dists, gmap_vpids, gmap_visited_masks,
                    global_act_labels, local_act_labels, compute_loss):
        """pretrain_cmt.py:218-262: node logits with visited / padded nodes at -inf, per-episode CE."""
        out = self.bert.forward_gmap(txt_ids, txt_lens, traj_view_img_fts, traj_view_dep_fts, traj_obj_img_fts,
                                     traj_loc_fts, traj_nav_types, traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens,
                                     traj_vpids, traj_cand_vpids, gmap_lens, gmap_step_ids, gmap_pos_fts, gmap_pair_dists,
                                     gmap_vpids, gmap_visited_masks)
        global_logits = out["global_logits"]
        labels = global_act_labels.to(global_logits.device)
        if compute_loss:
            return F.cross_entropy(global_logits, labels, reduction="none")
        return global_logits, labels


class PretrainTrainer:
    """Fused pre-training iteration (pretrain_src/pretrain_src/train_r2r.py:231-297): one task batch forward + backward
    through the step-level C calls with every parameter gradient accumulated straight into ONE flat fp32 buffer, one NCCL
    all-reduce of that buffer when data-parallel (the reference wraps the model in DDP, train_r2r.py:113-116; rank 0's
    parameters are broadcast first, as DDP does), and the fused AdamW over the whole flat parameter buffer with the
    reference's optimizer semantics: betas (0.9, 0.98), weight decay 0.01 except for biases and LayerNorm
    (optim/misc.py:14-20), global gradient-norm clipping at 5.0 (run_pt/r2r_pretrain_habitat.json ``grad_norm``,
    train_r2r.py:279-284), frozen parameters (``requires_grad = False``) left untouched."""

    NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")   # optim/misc.py:14

    def __init__(self, model: "B200PreTraining", lr=5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01, world_size=1,
                 grad_norm=5.0):
        self.model, self.m = model, model.bert
        self.lr, self.betas, self.eps, self.wd, self.world = lr, betas, eps, weight_decay, world_size
        self.grad_norm = grad_norm if (grad_norm is not None and grad_norm > 0) else 0.0
        m = self.m
        dev = m._flat.device
        if world_size > 1:
            import torch.distributed as dist
            with torch.no_grad():
                dist.broadcast(m._flat, src=0)
            m._cache_key = None
            if m._drop_base is None:
                m._drop_base = (torch.initial_seed() + 0x9E3779B97F4A7C15 * (dist.get_rank() + 1)) & 0xFFFFFFFFFFFFFFFF
        m._refresh_cache()
        n = m.layout.total
        m._direct_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flags = self.block_flags(m).to(dev)
        self.normsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.t = 0
        L = _L.lib()
        L.etp_adamw_step_ex.argtypes = [p_void, p_void, p_void, p_void, p_void, C.c_int64, C.c_float, C.c_float, C.c_float,
                                        C.c_float, C.c_float, i32, C.c_float, p_void, p_void, C.c_float, p_void]
        L.etp_grad_sumsq.argtypes = [p_void, C.c_int64, p_void, p_void, p_void]

    @classmethod
    def block_flags(cls, m):
        """One byte per 64 elements of the flat layout: bit 0 trainable, bit 1 weight decay (host tensor)."""
        assert m.layout.total % 64 == 0
        fl = torch.zeros(m.layout.total // 64, dtype=torch.uint8)
        for name, (off, numel, _) in m.layout.entries.items():
            assert off % 64 == 0
            if not m._pmap[name].requires_grad:
                continue
            v = 1 | (0 if any(nd in name for nd in cls.NO_DECAY) else 2)
            fl[off // 64:(off + numel + 63) // 64] = v
        return fl

    def step(self, batch, task):
        m = self.m
        m._direct_grad.zero_()
        loss = self.model(batch, task, compute_loss=True).mean()
        loss.backward()
        scale = 1.0
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(m._direct_grad, op=dist.ReduceOp.SUM)
            scale = 1.0 / self.world
        self.t += 1
        L = _L.lib()
        n = m.layout.total
        normsq = None
        if self.grad_norm > 0:
            self.normsq.zero_()
            _L._check(L.etp_grad_sumsq(_L.ptr(m._direct_grad), n, _L.ptr(self.flags), _L.ptr(self.normsq), _L.stream_ptr()),
                      "etp_grad_sumsq")
            normsq = self.normsq
        _L._check(L.etp_adamw_step_ex(_L.ptr(m._flat), _L.ptr(m._flat_bf16), _L.ptr(m._direct_grad), _L.ptr(self.exp_avg),
                                      _L.ptr(self.exp_avg_sq), n, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                      self.t, scale, _L.ptr(self.flags), _L.ptr(normsq), self.grad_norm, _L.stream_ptr()),
                  "etp_adamw_step_ex")
        m._bf16_fresh = True
        return loss
