// The 9-channel inpainting UNet and the DDIM loop on top of the engine's kernels.
// Structure follows the reference block for block so that state-dict keys map 1:1:
//   UNetModel.__init__/forward      ldm/modules/diffusionmodules/openaimodel.py:665-907
//   ResBlock._forward               openaimodel.py:255-275
//   SpatialTransformer / BasicTransformerBlock / CrossAttention / GEGLU   ldm/modules/attention.py:37-289
//   DDIMSampler.ddim_sampling / p_sample_ddim   ldm/models/diffusion/ddim.py:200-251, 323-375
#include "models.h"

namespace rfb {

static std::string J(const std::string& a, const std::string& b) { return a + b; }

static ResW build_res(Ctx& c, const std::string& p, int cin, int cout) {
  ResW r;
  r.cin = cin, r.cout = cout;
  r.g1 = c.pf(p + "in_layers.0.weight"), r.b1 = c.pf(p + "in_layers.0.bias");
  r.c1 = pack_conv(c, p + "in_layers.2.weight", p + "in_layers.2.bias");
  r.emb = lin32(c, p + "emb_layers.1.weight", p + "emb_layers.1.bias");
  r.g2 = c.pf(p + "out_layers.0.weight"), r.b2 = c.pf(p + "out_layers.0.bias");
  r.c2 = pack_conv(c, p + "out_layers.3.weight", p + "out_layers.3.bias");
  r.skip = cin != cout;
  if (r.skip) r.skipw = pack_conv(c, p + "skip_connection.weight", p + "skip_connection.bias");
  RFB_CHECK(r.c1.cin == cin && r.c1.cout == cout, "ResBlock weight shape mismatch at " + p);
  return r;
}

static STW build_st(Ctx& c, const std::string& p, int ch, int heads, int ctx_dim) {
  STW s;
  s.c = ch, s.heads = heads, s.d = ch / heads, s.ctx_dim = ctx_dim;
  const std::string b = p + "transformer_blocks.0.";
  s.gn_g = c.pf(p + "norm.weight"), s.gn_b = c.pf(p + "norm.bias");
  s.proj_in = pack_conv(c, p + "proj_in.weight", p + "proj_in.bias");
  s.proj_in_w32 = c.pf(p + "proj_in.weight"), s.proj_in_b = c.pf(p + "proj_in.bias");
  s.proj_out = pack_conv(c, p + "proj_out.weight", p + "proj_out.bias");
  s.ln1g = c.pf(b + "norm1.weight"), s.ln1b = c.pf(b + "norm1.bias");
  s.ln2g = c.pf(b + "norm2.weight"), s.ln2b = c.pf(b + "norm2.bias");
  s.ln3g = c.pf(b + "norm3.weight"), s.ln3b = c.pf(b + "norm3.bias");
  s.qkv = pack_linear_rows(c, {b + "attn1.to_q.weight", b + "attn1.to_k.weight", b + "attn1.to_v.weight"});
  s.o1 = pack_linear(c, b + "attn1.to_out.0.weight", b + "attn1.to_out.0.bias");
  s.v2 = lin32(c, b + "attn2.to_v.weight", "");
  s.k2 = lin32(c, b + "attn2.to_k.weight", "");
  s.o2 = lin32(c, b + "attn2.to_out.0.weight", b + "attn2.to_out.0.bias");
  // general (context length > 1) cross-attention operands
  s.q2 = pack_linear(c, b + "attn2.to_q.weight", "");
  s.kv2 = pack_linear_rows(c, {b + "attn2.to_k.weight", b + "attn2.to_v.weight"});
  s.o2h = pack_linear(c, b + "attn2.to_out.0.weight", b + "attn2.to_out.0.bias");
  s.ff_bn = 256;
  while ((8 * ch) % s.ff_bn) s.ff_bn /= 2;
  s.ff1 = pack_geglu(c, b + "ff.net.0.proj.weight", b + "ff.net.0.proj.bias", s.ff_bn);
  s.ff2 = pack_linear(c, b + "ff.net.2.weight", b + "ff.net.2.bias");
  return s;
}

UNet* build_unet(Ctx& c, const std::string& pfx, const UNetCfg& cfg) {
  UNet* u = new UNet();
  u->cfg = cfg;
  u->pfx = pfx;
  const int mc = cfg.model_channels;
  u->te0 = lin32(c, pfx + "time_embed.0.weight", pfx + "time_embed.0.bias");
  u->te2 = lin32(c, pfx + "time_embed.2.weight", pfx + "time_embed.2.bias");
  auto has_attn = [&](int ds) {
    for (int a : cfg.attention_resolutions)
      if (a == ds) return true;
    return false;
  };
  // input blocks
  {
    UOp op;
    op.kind = OP_CONV_IN;
    op.conv = pack_conv(c, pfx + "input_blocks.0.0.weight", pfx + "input_blocks.0.0.bias");
    u->inp.push_back({op});
  }
  std::vector<int> chans = {mc};
  int ch = mc, ds = 1, idx = 1;
  const int nlev = (int)cfg.channel_mult.size();
  for (int level = 0; level < nlev; ++level) {
    const int m = cfg.channel_mult[level];
    for (int i = 0; i < cfg.num_res_blocks; ++i) {
      std::vector<UOp> ops;
      const std::string bp = pfx + "input_blocks." + std::to_string(idx) + ".";
      UOp r;
      r.kind = OP_RES;
      r.res = build_res(c, bp + "0.", ch, m * mc);
      ops.push_back(r);
      ch = m * mc;
      if (has_attn(ds)) {
        UOp a;
        a.kind = OP_ATTN;
        a.st = build_st(c, bp + "1.", ch, cfg.num_heads, cfg.context_dim);
        ops.push_back(a);
      }
      u->inp.push_back(ops);
      chans.push_back(ch);
      ++idx;
    }
    if (level != nlev - 1) {
      UOp d;
      d.kind = OP_DOWN;
      const std::string bp = pfx + "input_blocks." + std::to_string(idx) + ".0.op.";
      d.conv = pack_conv(c, bp + "weight", bp + "bias");
      u->inp.push_back({d});
      chans.push_back(ch);
      ds *= 2;
      ++idx;
    }
  }
  // middle
  {
    UOp r1, a, r2;
    r1.kind = OP_RES, r1.res = build_res(c, pfx + "middle_block.0.", ch, ch);
    a.kind = OP_ATTN, a.st = build_st(c, pfx + "middle_block.1.", ch, cfg.num_heads, cfg.context_dim);
    r2.kind = OP_RES, r2.res = build_res(c, pfx + "middle_block.2.", ch, ch);
    u->mid = {r1, a, r2};
  }
  // output blocks
  idx = 0;
  for (int level = nlev - 1; level >= 0; --level) {
    const int m = cfg.channel_mult[level];
    for (int i = 0; i <= cfg.num_res_blocks; ++i) {
      const int ich = chans.back();
      chans.pop_back();
      std::vector<UOp> ops;
      const std::string bp = pfx + "output_blocks." + std::to_string(idx) + ".";
      int j = 0;
      UOp r;
      r.kind = OP_RES;
      r.res = build_res(c, bp + std::to_string(j++) + ".", ch + ich, mc * m);
      ops.push_back(r);
      ch = mc * m;
      if (has_attn(ds)) {
        UOp a;
        a.kind = OP_ATTN;
        a.st = build_st(c, bp + std::to_string(j++) + ".", ch, cfg.num_heads, cfg.context_dim);
        ops.push_back(a);
      }
      if (level && i == cfg.num_res_blocks) {
        UOp up;
        up.kind = OP_UP;
        const std::string cp = bp + std::to_string(j++) + ".conv.";
        up.conv = pack_upconv(c, cp + "weight", cp + "bias");  // nearest-2x folded into the conv (engine.cu)
        ops.push_back(up);
        ds /= 2;
      }
      u->out.push_back(ops);
      ++idx;
    }
  }
  // all 22 ResBlock emb_layers (Linear(1280 -> Cout) on SiLU(emb)) share their input: concatenate them so that
  // one launch per forward produces every block's time-embedding row
  {
    std::vector<ResW*> rs;
    auto visit = [&](std::vector<UOp>& ops) {
      for (UOp& op : ops)
        if (op.kind == OP_RES) rs.push_back(&op.res);
    };
    for (auto& ops : u->inp) visit(ops);
    visit(u->mid);
    for (auto& ops : u->out) visit(ops);
    int total = 0;
    for (ResW* r : rs) r->emb_off = total, total += r->emb.out;
    const int in = rs.empty() ? 0 : rs[0]->emb.in;
    float* w = (float*)c.dmalloc((size_t)total * in * sizeof(float));
    float* b = (float*)c.dmalloc((size_t)total * sizeof(float));
    for (ResW* r : rs) {
      CUDA_OK(cudaMemcpyAsync(w + (size_t)r->emb_off * in, r->emb.w, (size_t)r->emb.out * in * sizeof(float),
                              cudaMemcpyDeviceToDevice, c.stream));
      CUDA_OK(cudaMemcpyAsync(b + r->emb_off, r->emb.b, (size_t)r->emb.out * sizeof(float), cudaMemcpyDeviceToDevice,
                              c.stream));
    }
    u->emb_cat.w = w, u->emb_cat.b = b, u->emb_cat.in = in, u->emb_cat.out = total;
  }
  u->out_g = c.pf(pfx + "out.0.weight"), u->out_b = c.pf(pfx + "out.0.bias");
  u->out_taps = pack_out_taps(c, pfx + "out.2.weight");
  u->out_bias = c.pf(pfx + "out.2.bias");
  CUDA_OK(cudaStreamSynchronize(c.stream));
  return u;
}

// ---------------------------------------------------------------------------------------------- forward
// x2.p != nullptr: the block's input is torch.cat([x, x2], 1) (openaimodel.py:897-899), never materialised: GroupNorm
// reads both tensors, the 1x1 skip_connection conv runs one K loop over two TMA descriptors.  x2 may hold half the
// samples of x (the conv_in output shared by the two CFG halves).
static Tens run_res(Ctx& c, const ResW& r, const Tens& x, const float* emb_all, int emb_rows, int emb_ld,
                    const Tens& x2 = Tens()) {
  // emb_out = Linear(SiLU(emb)) [openaimodel.py:264] was computed for all blocks at once (emb_all: [rows, emb_ld]);
  // emb_rows == 1 when every sample shares the timestep
  Tens h = groupnorm2(c, x, x2, r.g1, r.b1, 1e-5f, true);
  Epi e1;
  e1.rowvec = emb_all + r.emb_off, e1.ldv = emb_rows == 1 ? 0 : emb_ld;
  e1.want_stats = true;  // h1 feeds GroupNorm: its statistics come out of this conv's epilogue
  Tens h1 = conv3x3_t(c, h, r.c1, e1);
  Tens h2 = groupnorm(c, h1, r.g2, r.b2, 1e-5f, true);
  Tens skip = x;
  if (x2.p) {
    RFB_CHECK(r.skip && r.skipw.ksz == 1, "a ResBlock over a concatenated input has a 1x1 skip_connection");
    skip = c.new_tens(x.n, x.h, x.w, r.cout);
    Epi es;
    es.bias = r.skipw.b;
    gemm2(c, x.p, x.c, x.c, x2.p, x2.c, x2.c, x2.n != x.n ? x2.rows() : 0, x.rows(), r.skipw.w, r.skipw.kp, r.cout, skip.p,
          r.cout, es);
  } else if (r.skip) {
    skip = conv3x3_t(c, x, r.skipw, Epi(), 1, 0, 0, 0, 0);
  }
  Epi e2;
  e2.res = skip.p, e2.ldr = skip.c;
  e2.want_stats = true;  // the block output feeds the next block's GroupNorm
  return conv3x3_t(c, h2, r.c2, e2);
}

// attn2 with a single context token: softmax over one key is exactly 1, so the block adds
// to_out(to_v(ctx[n])) to every token (attention.py:206-221); the vector depends on the context only.
static float* cross_vec(Ctx& c, const STW& s, const float* ctx, int N) {
  float* v = c.alloc_t<float>((size_t)N * s.c);
  float* vec = c.alloc_t<float>((size_t)N * s.c);
  linear_small(c, ctx, s.ctx_dim, N, s.v2, v, s.c, 0, 0);
  linear_small(c, v, s.c, N, s.o2, vec, s.c, 0, 0);
  return vec;
}

// share_halves: x holds ONE copy of the N/2 distinct samples of a CFG batch (UNetAux::cfg_dup); GroupNorm, proj_in,
// LN1, the QKV projection and the self-attention do not see the context and run once, attn1's out-projection is issued
// per half with that half's cross-attention vectors, and the block continues with all N samples.
static Tens run_st(Ctx& c, const STW& s, const Tens& x, const float* ctx, int T, int N, const float* vec_pre,
                   bool share_halves = false) {
  const int C = s.c;
  const int L = x.h * x.w;
  Tens h;
  if (c.gn_fold && c.gn_epi_stats && x.stats && L >= 1024 && L % 128 == 0) {
    // norm (no activation) + proj_in as ONE GEMM over the raw tensor with per-sample folded weights
    h = conv1x1_gn_folded(c, x, s.gn_g, s.gn_b, 1e-6f, s.proj_in_w32, s.proj_in_b, C);
  } else {
    Tens xn = groupnorm(c, x, s.gn_g, s.gn_b, 1e-6f, false);
    h = conv3x3_t(c, xn, s.proj_in, Epi(), 1, 0, 0, 0, 0);
  }
  // --- attn1 (self attention) [attention.py:240]
  Tens n1 = layernorm(c, h, s.ln1g, s.ln1b, 1e-5f);
  Tens qkv = linear_t(c, n1, s.qkv, Epi());
  Tens a = c.new_tens(x.n, x.h, x.w, C);
  attention(c, qkv.p, 3 * C, x.n, L, s.heads, s.d, a.p, C, 1.0f / sqrtf((float)s.d), 0, C, 2 * C);
  Tens h1;
  long long xres_mod = 0;  // residual of proj_out: x itself; shared halves read it modulo its rows
  if (share_halves) {
    RFB_CHECK(T == 1 && N == 2 * x.n, "shared CFG halves need the single-token context path");
    const float* vec = vec_pre ? vec_pre : cross_vec(c, s, ctx, N);
    h1 = c.new_tens(N, x.h, x.w, C);
    const long long Mh = x.rows();
    for (int half = 0; half < 2; ++half) {
      Epi e;
      e.bias = s.o1.b, e.res = h.p, e.ldr = C, e.rowvec = vec + (size_t)half * x.n * C, e.ldv = C, e.rows_per_vec = L;
      gemm(c, a.p, C, Mh, C, s.o1.w, s.o1.kp, s.o1.out, h1.p + (size_t)half * Mh * C, C, e);
    }
    xres_mod = x.rows();
  } else if (T == 1) {
    // --- attn2 (degenerate, see cross_vec) folded into attn1's out-projection epilogue
    const float* vec = vec_pre ? vec_pre : cross_vec(c, s, ctx, N);
    Epi e;
    e.res = h.p, e.ldr = C, e.rowvec = vec, e.ldv = C, e.rows_per_vec = L;
    h1 = linear_t(c, a, s.o1, e);
  } else {
    Epi e;
    e.res = h.p, e.ldr = C;
    Tens h0 = linear_t(c, a, s.o1, e);
    h1 = cross_attention_general(c, s, h0, ctx, T, N);
  }
  // --- feed-forward (GEGLU) [attention.py:37-64,242]
  Tens n3 = layernorm(c, h1, s.ln3g, s.ln3b, 1e-5f);
  Epi eg;
  eg.geglu = 1;
  Tens ff = linear_t(c, n3, s.ff1, eg);
  Epi e2;
  e2.res = h1.p, e2.ldr = C;
  Tens h2 = linear_t(c, ff, s.ff2, e2);
  Epi e3;
  e3.res = x.p, e3.ldr = C, e3.res_mod = xres_mod;
  e3.want_stats = true;
  return conv3x3_t(c, h2, s.proj_out, e3, 1, 0, 0, 0, 0);
}

// General cross-attention for context length T > 1 (stack_feat configs, ddpm.py:1027-1030; attention.py:179-221):
//   x + to_out(softmax(to_q(LN2 x) to_k(ctx)^T / sqrt d) to_v(ctx)).  The context projections are tiny (T rows per
// sample, fp32 GEMV path); the T-wide softmax runs on CUDA cores; q and the out-projection use the tensor-core GEMM.
Tens cross_attention_general(Ctx& c, const STW& s, const Tens& x, const float* ctx, int T, int N) {
  const int C = s.c, L = x.h * x.w;
  Tens n2 = layernorm(c, x, s.ln2g, s.ln2b, 1e-5f);
  Tens q = linear_t(c, n2, s.q2, Epi());
  float* kc = c.alloc_t<float>((size_t)N * T * C);
  float* vc = c.alloc_t<float>((size_t)N * T * C);
  linear_small(c, ctx, s.ctx_dim, N * T, s.k2, kc, C, 0, 0);
  linear_small(c, ctx, s.ctx_dim, N * T, s.v2, vc, C, 0, 0);
  Tens a = c.new_tens(x.n, x.h, x.w, C);
  cross_attn_small(c, q.p, kc, vc, a.p, N, L, T, C, s.heads);
  Epi e;
  e.res = x.p, e.ldr = C;
  return linear_t(c, a, s.o2h, e);
}

struct RunState {
  const float* emb;  // every ResBlock's emb_layers output: [emb_rows, emb_ld]
  int emb_rows, emb_ld;
  const float* ctx;
  int T, N;
  const UNetAux* aux;
  int st_idx;
};

static Epi stats_epi() {  // plain epilogue that also leaves GroupNorm partial statistics with the output
  Epi e;
  e.want_stats = true;
  return e;
}
// cat.p != nullptr: the first op (a ResBlock) consumes torch.cat([h, cat], 1)
static Tens run_ops(Ctx& c, const std::vector<UOp>& ops, Tens h, RunState& rs, Tens cat = Tens()) {
  for (const UOp& op : ops) {
    switch (op.kind) {
      case OP_RES:
        h = run_res(c, op.res, h, rs.emb, rs.emb_rows, rs.emb_ld, cat);
        cat = Tens();
        break;
      case OP_ATTN: {
        const float* pre = (rs.aux && rs.st_idx < (int)rs.aux->crossvec.size()) ? rs.aux->crossvec[rs.st_idx] : nullptr;
        h = run_st(c, op.st, h, rs.ctx, rs.T, rs.N, pre);
        ++rs.st_idx;
      } break;
      case OP_DOWN: h = conv3x3_t(c, h, op.conv, stats_epi(), 2, 1, 1, 1, 1); break;
      case OP_UP: h = upconv3x3_t(c, h, op.conv, stats_epi()); break;
      case OP_CONV_IN: h = conv3x3_t(c, h, op.conv, stats_epi()); break;
    }
  }
  return h;
}

// x9: [N,9,L,L] fp32 NCHW, t: [N] int64, ctx: [N,T,768] fp32 (all device) -> eps [N,4,L,L] fp32 NCHW
std::vector<const float*> unet_cross_vectors(Ctx& c, UNet& u, const float* ctx, int N, int T) {
  std::vector<const float*> out;
  if (T != 1) return out;
  auto visit = [&](const std::vector<UOp>& ops) {
    for (const UOp& op : ops)
      if (op.kind == OP_ATTN) out.push_back(cross_vec(c, op.st, ctx, N));
  };
  for (auto& ops : u.inp) visit(ops);
  visit(u.mid);
  for (auto& ops : u.out) visit(ops);
  return out;
}

// timestep_embedding -> time_embed MLP (openaimodel.py:874-875) -> every ResBlock's emb_layers (openaimodel.py:264) for R
// timesteps at once.  The rows do not depend on each other or on R (fixed K order per row): bitwise the per-call values.
float* unet_time_embeddings(Ctx& c, UNet& u, const long long* t, int R) {
  const int mc = u.cfg.model_channels;
  float* emb_all = c.alloc_t<float>((size_t)R * u.emb_cat.out);
  const size_t mk = c.mark();
  float* temb = c.alloc_t<float>((size_t)R * mc);
  float* e1 = c.alloc_t<float>((size_t)R * 4 * mc);
  float* emb = c.alloc_t<float>((size_t)R * 4 * mc);
  timestep_embedding(c, t, temb, R, mc);
  linear_small(c, temb, mc, R, u.te0, e1, 4 * mc, 0, /*silu*/ 1);
  linear_small(c, e1, 4 * mc, R, u.te2, emb, 4 * mc, 0, 0);
  linear_small(c, emb, 4 * mc, R, u.emb_cat, emb_all, u.emb_cat.out, /*act_in=silu*/ 1, 0);
  c.release(mk);
  return emb_all;
}

void unet_forward(Ctx& c, UNet& u, const float* x9, const long long* t, const float* ctx, int N, int L, int T,
                  float* eps, const UNetAux* aux) {
  const size_t mk = c.mark();
  // one time-embedding row is enough when all samples share the timestep
  const int er = (aux && aux->uniform_t) ? 1 : N;
  const float* emb_all = (aux && aux->uniform_t && aux->emb_all) ? aux->emb_all : unet_time_embeddings(c, u, t, er);
  RunState rs{emb_all, er, u.emb_cat.out, ctx, T, N, aux, 0};
  std::vector<Tens> hs;
  Tens h;
  size_t first = 0;
  const bool share = aux && aux->cfg_dup && c.cfg_share && T == 1 && N % 2 == 0 && u.inp.size() >= 2 &&
                     u.inp[0].size() == 1 && u.inp[0][0].kind == OP_CONV_IN && u.inp[1].size() == 2 &&
                     u.inp[1][0].kind == OP_RES && u.inp[1][1].kind == OP_ATTN;
  if (share) {
    // CFG batch: the two halves only differ in the context, which first enters in attn2 of input_blocks.1
    Tens x8 = from_nchw_f32(c, x9, N / 2, u.cfg.in_channels, L, L, u.cfg.in_channels);
    Tens h0 = conv3x3_t(c, x8, u.inp[0][0].conv, stats_epi());
    hs.push_back(h0);  // N/2 samples: the consumers (last output block) read it modulo its sample count
    Tens r = run_res(c, u.inp[1][0].res, h0, rs.emb, rs.emb_rows, rs.emb_ld);
    const float* pre = !aux->crossvec.empty() ? aux->crossvec[0] : nullptr;
    h = run_st(c, u.inp[1][1].st, r, ctx, T, N, pre, true);
    rs.st_idx = 1;
    hs.push_back(h);
    first = 2;
  } else {
    h = from_nchw_f32(c, x9, N, u.cfg.in_channels, L, L, u.cfg.in_channels);
  }
  for (size_t bi = first; bi < u.inp.size(); ++bi) {
    h = run_ops(c, u.inp[bi], h, rs);
    hs.push_back(h);
  }
  h = run_ops(c, u.mid, h, rs);
  for (auto& ops : u.out) {
    RFB_CHECK(!ops.empty() && ops[0].kind == OP_RES, "output blocks start with a ResBlock");
    const Tens skip = hs.back();
    hs.pop_back();
    h = run_ops(c, ops, h, rs, skip);
  }
  Tens hn = groupnorm(c, h, u.out_g, u.out_b, 1e-5f, true);
  // out.2 (3x3, 320 -> 4): one [N*L*L, 320] x [320, 36] tap GEMM, finished by a 9-tap gather (elem.cuh: eps_gather)
  RFB_CHECK(u.cfg.out_channels == 4, "the tap formulation of the output conv is written for 4 output channels");
  const long long M = hn.rows();
  float* taps = (aux && aux->taps_out) ? aux->taps_out : c.alloc_t<float>((size_t)M * 36);
  Epi e;
  e.out32 = taps, e.o32_sn = 36, e.o32_sp = 0, e.o32_sc = 1, e.o32_rpn = 1;
  gemm(c, hn.p, hn.c, M, hn.c, u.out_taps.w, u.out_taps.kp, 36, nullptr, 0, e);
  if (!(aux && aux->taps_out)) eps_from_taps(c, taps, u.out_bias, eps, N, L);
  c.release(mk);
}

// ---------------------------------------------------------------------------------------------- DDIM loop
// DDIMSampler.ddim_sampling + p_sample_ddim (ddim.py:200-251, 323-375).  The loop body works on arena-resident buffers
// only (inputs are staged in, results staged out), so that the WHOLE S-step loop -- every kernel of every step with its
// own schedule scalars -- can be captured once into a CUDA graph and replayed by later calls with the same shape /
// schedule: arena addresses are a deterministic function of the arena mark at entry, which is part of the
// cache key.  First call with a key: eager; second call: capture + instantiate + launch; afterwards: one cudaGraphLaunch.
struct DdimBufs {
  float *xa, *xb, *p0, *x9, *taps, *ctx, *z, *mask, *o_x0, *o_ix, *o_ip;
  long long *ts, *ts1;  // timesteps per (step, sample) / per step
};

static int ddim_body(Ctx& c, UNet& u, DdimBufs b, int B, int L, int T, const DdimSchedule& s, float scale, bool cfg,
                     const float* noise, int log_every_t) {
  const long long HW = (long long)L * L, cnt = (long long)B * 4 * HW;
  const int dup = cfg ? 2 : 1, N = dup * B;
  // step-invariant terms, computed once per sampling run: the cross-attention vectors depend on
  // the context only; every sample shares the step's timestep, so one time-embedding row serves the batch.
  UNetAux aux;
  aux.uniform_t = 1;
  aux.cfg_dup = cfg ? 1 : 0;
  aux.crossvec = unet_cross_vectors(c, u, b.ctx, N, T);
  aux.taps_out = b.taps;
  // ... and so does the whole time-embedding path: emb_layers(time_embed(t)) for ALL steps in four launches
  // (per step it streamed the 115 MB of fp32 emb_layers weights for a single row)
  const float* emb_table = unet_time_embeddings(c, u, b.ts1, s.n);
  int n_inter = 0;
  float *xa = b.xa, *xb = b.xb;
  for (int i = 0; i < s.n; ++i) {
    const int index = s.n - 1 - i;
    aux.emb_all = emb_table + (size_t)index * u.emb_cat.out;
    concat9(c, xa, b.z, b.mask, b.x9, B, (int)HW, dup);
    unet_forward(c, u, b.x9, b.ts + (size_t)index * N, b.ctx, N, L, T, nullptr, &aux);
    // CFG combine + DDIM update fused with the end of the UNet's output convolution: eps never reaches HBM
    taps_cfg_ddim_update(c, xa, b.taps, u.out_bias, noise ? noise + (size_t)i * cnt : nullptr, xb, b.p0, B, L, scale,
                         s.a_t[index], s.a_prev[index], s.sigma[index], s.sqrt_one_minus_a[index], cfg ? 1 : 0);
    std::swap(xa, xb);
    if (log_every_t > 0 && (index % log_every_t == 0 || index == s.n - 1)) {  // ddim.py:247-249
      CUDA_OK(cudaMemcpyAsync(b.o_ix + (size_t)n_inter * cnt, xa, cnt * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
      CUDA_OK(cudaMemcpyAsync(b.o_ip + (size_t)n_inter * cnt, b.p0, cnt * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
      ++n_inter;
    }
  }
  CUDA_OK(cudaMemcpyAsync(b.o_x0, xa, cnt * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  return n_inter;
}

void ddim_sample(Ctx& c, UNet& u, const float* x_T, const float* z_inpaint, const float* mask, const float* cond,
                 const float* uncond, int B, int L, int T, const DdimSchedule& s, float scale, const float* noise,
                 float* x0_out, float* inter_x, float* inter_p0, int log_every_t) {
  const size_t mk = c.mark();
  const long long HW = (long long)L * L, cnt = (long long)B * 4 * HW;
  const bool cfg = uncond != nullptr && scale != 1.0f;  // ddim.py:335
  const int dup = cfg ? 2 : 1, N = dup * B;
  int K = 0;  // logged intermediates
  for (int index = 0; index < s.n; ++index)
    if (log_every_t > 0 && (index % log_every_t == 0 || index == s.n - 1)) ++K;
  DdimBufs b;
  b.xa = c.alloc_t<float>(cnt), b.xb = c.alloc_t<float>(cnt), b.p0 = c.alloc_t<float>(cnt);
  b.x9 = c.alloc_t<float>((size_t)N * 9 * HW), b.taps = c.alloc_t<float>((size_t)N * HW * 36);
  b.ctx = c.alloc_t<float>((size_t)N * T * 768);
  b.z = c.alloc_t<float>(cnt), b.mask = c.alloc_t<float>((size_t)B * HW);
  b.o_x0 = c.alloc_t<float>(cnt);
  b.o_ix = c.alloc_t<float>((size_t)std::max(K, 1) * cnt), b.o_ip = c.alloc_t<float>((size_t)std::max(K, 1) * cnt);
  b.ts = c.alloc_t<long long>((size_t)s.n * N + s.n);
  b.ts1 = b.ts + (size_t)s.n * N;
  // ---- prologue on the caller's stream: stage the inputs into the arena
  {
    std::vector<long long> h((size_t)s.n * N + s.n);
    for (int i = 0; i < s.n; ++i) {
      for (int j = 0; j < N; ++j) h[(size_t)i * N + j] = s.timesteps[i];
      h[(size_t)s.n * N + i] = s.timesteps[i];
    }
    CUDA_OK(cudaMemcpyAsync(b.ts, h.data(), h.size() * sizeof(long long), cudaMemcpyHostToDevice, c.stream));
    CUDA_OK(cudaStreamSynchronize(c.stream));  // h goes out of scope
  }
  const size_t cb = (size_t)B * T * 768 * sizeof(float);
  if (cfg) {  // c_in = cat([uc, c])  ddim.py:344
    CUDA_OK(cudaMemcpyAsync(b.ctx, uncond, cb, cudaMemcpyDeviceToDevice, c.stream));
    CUDA_OK(cudaMemcpyAsync(b.ctx + (size_t)B * T * 768, cond, cb, cudaMemcpyDeviceToDevice, c.stream));
  } else {
    CUDA_OK(cudaMemcpyAsync(b.ctx, cond, cb, cudaMemcpyDeviceToDevice, c.stream));
  }
  CUDA_OK(cudaMemcpyAsync(b.xa, x_T, cnt * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  CUDA_OK(cudaMemcpyAsync(b.z, z_inpaint, cnt * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  CUDA_OK(cudaMemcpyAsync(b.mask, mask, (size_t)B * HW * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));

  // ---- the loop: eager, or one graph launch
  bool done = false;
  if (c.use_graph && !c.profile && !c.gemm_debug && noise == nullptr) {
    // key: everything the captured launches depend on besides the staged buffer contents
    std::string key = "ddim";
    auto add = [&](const void* p, size_t n) { key.append(reinterpret_cast<const char*>(p), n); };
    const long long dims[8] = {B, L, T, s.n, log_every_t, cfg ? 1 : 0, (long long)mk, c.graph_epoch};
    add(dims, sizeof(dims));
    add(&scale, sizeof(scale));
    add(s.timesteps.data(), s.timesteps.size() * sizeof(long long));
    add(s.a_t.data(), s.a_t.size() * sizeof(float));
    add(s.a_prev.data(), s.a_prev.size() * sizeof(float));
    add(s.sigma.data(), s.sigma.size() * sizeof(float));
    add(s.sqrt_one_minus_a.data(), s.sqrt_one_minus_a.size() * sizeof(float));
    if (c.graphs.size() >= 16 && !c.graphs.count(key)) {  // bounded cache: a server cycling through many shapes starts over
      const long long epoch = c.graph_epoch;
      c.drop_graphs();
      c.graph_epoch = epoch;  // same options / weights: keys stay valid
    }
    Ctx::GraphEntry& ge = c.graphs[key];
    ++ge.seen;
    if (!ge.exec && !ge.failed && ge.seen >= 2) {
      if (!c.gstream) CUDA_OK(cudaStreamCreateWithFlags(&c.gstream, cudaStreamNonBlocking));
      cudaStream_t user = c.stream;
      const long long l0 = c.launches;
      const size_t mk2 = c.mark();
      c.stream = c.gstream;
      cudaGraph_t graph = nullptr;
      try {
        CUDA_OK(cudaStreamBeginCapture(c.gstream, cudaStreamCaptureModeRelaxed));
        ddim_body(c, u, b, B, L, T, s, scale, cfg, nullptr, log_every_t);
        CUDA_OK(cudaStreamEndCapture(c.gstream, &graph));
        CUDA_OK(cudaGraphInstantiate(&ge.exec, graph, 0));
      } catch (const std::exception&) {
        cudaGraph_t junk = nullptr;
        cudaStreamEndCapture(c.gstream, &junk);  // leave capture mode whatever happened
        if (junk) cudaGraphDestroy(junk);
        cudaGetLastError();
        ge.exec = nullptr, ge.failed = true;
      }
      if (graph) cudaGraphDestroy(graph);
      c.stream = user;
      c.release(mk2);
      ge.launches = c.launches - l0;
      c.launches = l0;
    }
    if (ge.exec) {
      CUDA_OK(cudaGraphLaunch(ge.exec, c.stream));
      c.launches += ge.launches;
      c.graph_replays++;
      done = true;
    }
  }
  if (!done) ddim_body(c, u, b, B, L, T, s, scale, cfg, noise, log_every_t);

  // ---- epilogue: results to the caller's buffers
  CUDA_OK(cudaMemcpyAsync(x0_out, b.o_x0, cnt * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  if (K > 0 && inter_x)
    CUDA_OK(cudaMemcpyAsync(inter_x, b.o_ix, (size_t)K * cnt * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  if (K > 0 && inter_p0)
    CUDA_OK(cudaMemcpyAsync(inter_p0, b.o_ip, (size_t)K * cnt * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  c.release(mk);
}

// ---------------------------------------------------------------------------------------------- PLMS loop
// PLMSSampler.plms_sampling / p_sample_plms with test_model_kwargs (ldm/models/diffusion/plms.py:116-242), eta = 0.
void plms_sample(Ctx& c, UNet& u, const float* x_T, const float* z_inpaint, const float* mask, const float* cond,
                 const float* uncond, int B, int L, int T, const DdimSchedule& s, float scale, float* x0_out,
                 float* inter_x, float* inter_p0, int log_every_t) {
  const size_t mk = c.mark();
  const long long HW = (long long)L * L, cnt = (long long)B * 4 * HW;
  const bool cfg = uncond != nullptr && scale != 1.0f;  // plms.py:179
  const int dup = cfg ? 2 : 1, N = dup * B;
  float* xa = c.alloc_t<float>(cnt);
  float* xb = c.alloc_t<float>(cnt);
  float* p0 = c.alloc_t<float>(cnt);
  float* e_next = c.alloc_t<float>(cnt);
  float* e_prime = c.alloc_t<float>(cnt);
  float* ring[4];
  for (int i = 0; i < 4; ++i) ring[i] = c.alloc_t<float>(cnt);
  float* x9 = c.alloc_t<float>((size_t)N * 9 * HW);
  float* eps2 = c.alloc_t<float>((size_t)N * 4 * HW);
  float* ctx = c.alloc_t<float>((size_t)N * T * 768);
  long long* ts = c.alloc_t<long long>((size_t)s.n * N);
  {
    std::vector<long long> h((size_t)s.n * N);
    for (int i = 0; i < s.n; ++i)
      for (int j = 0; j < N; ++j) h[(size_t)i * N + j] = s.timesteps[i];
    CUDA_OK(cudaMemcpyAsync(ts, h.data(), h.size() * sizeof(long long), cudaMemcpyHostToDevice, c.stream));
    CUDA_OK(cudaStreamSynchronize(c.stream));
  }
  const size_t cb = (size_t)B * T * 768 * sizeof(float);
  if (cfg) {
    CUDA_OK(cudaMemcpyAsync(ctx, uncond, cb, cudaMemcpyDeviceToDevice, c.stream));
    CUDA_OK(cudaMemcpyAsync(ctx + (size_t)B * T * 768, cond, cb, cudaMemcpyDeviceToDevice, c.stream));
  } else {
    CUDA_OK(cudaMemcpyAsync(ctx, cond, cb, cudaMemcpyDeviceToDevice, c.stream));
  }
  CUDA_OK(cudaMemcpyAsync(xa, x_T, cnt * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  UNetAux aux;
  aux.uniform_t = 1;
  aux.cfg_dup = cfg ? 1 : 0;
  aux.crossvec = unet_cross_vectors(c, u, ctx, N, T);
  // get_model_output (plms.py:178-192): e = eps(x, t) with classifier-free guidance
  auto model_output = [&](const float* x, int index, float* e) {
    concat9(c, x, z_inpaint, mask, x9, B, (int)HW, dup);
    unet_forward(c, u, x9, ts + (size_t)index * N, ctx, N, L, T, cfg ? eps2 : e, &aux);
    if (cfg) cfg_combine(c, eps2, e, cnt, scale);
  };
  // get_x_prev_and_pred_x0 (plms.py:199-217) with sigma = 0
  auto step_from = [&](const float* x, const float* e, int index, float* x_prev, float* pred) {
    cfg_ddim_update(c, x, e, nullptr, x_prev, pred, cnt, 1.0f, s.a_t[index], s.a_prev[index], s.sigma[index],
                    s.sqrt_one_minus_a[index], 0);
  };
  int n_inter = 0;
  for (int i = 0; i < s.n; ++i) {
    const int index = s.n - 1 - i;
    const int index_next = s.n - 1 - std::min(i + 1, s.n - 1);  // time_range[min(i + 1, len - 1)], plms.py:146
    float* e_t = ring[i & 3];
    model_output(xa, index, e_t);
    const int order = std::min(i, 3);
    if (order == 0) {  // pseudo improved Euler: second model evaluation at the predicted x_prev
      step_from(xa, e_t, index, xb, nullptr);
      model_output(xb, index_next, e_next);
    }
    plms_combine(c, e_t, ring[(i + 3) & 3], ring[(i + 2) & 3], ring[(i + 1) & 3], e_next, e_prime, cnt, order);
    step_from(xa, e_prime, index, xb, p0);
    std::swap(xa, xb);
    if (log_every_t > 0 && (index % log_every_t == 0 || index == s.n - 1)) {  // plms.py:168-170
      if (inter_x)
        CUDA_OK(cudaMemcpyAsync(inter_x + (size_t)n_inter * cnt, xa, cnt * sizeof(float), cudaMemcpyDeviceToDevice,
                                c.stream));
      if (inter_p0)
        CUDA_OK(cudaMemcpyAsync(inter_p0 + (size_t)n_inter * cnt, p0, cnt * sizeof(float), cudaMemcpyDeviceToDevice,
                                c.stream));
      ++n_inter;
    }
  }
  CUDA_OK(cudaMemcpyAsync(x0_out, xa, cnt * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  c.release(mk);
}

}  // namespace rfb
