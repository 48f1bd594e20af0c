// cse_model_* : the member-level entry points of the C ABI (include/cse.h).  A caller that is not Python builds a
// member with   cse_model_create(arch, T, H, W, classes, dtype, max_batch)  ->  cse_model_set_weight(layer, tensor,
// host fp32 in Keras layout)  ->  cse_model_finalize  ->  cse_model_forward   and never sees tiles, packings or
// buffer offsets: graph construction (model_graph.h) and lowering (this file + model_lower.h: fusion, engine and tile
// choice, BN folding, weight re-packing, liveness-based buffer planning) run inside the library.  Replaces
// evaluate_load_model + model.predict_generator of the reference (train.py:1712-1772, evaluate_ensemble.py:1053-1056).
#include <stdarg.h>

#include <functional>

#include "common.cuh"
#include "model_lower.h"

namespace cse {
namespace mdl {

class Lowerer {
 public:
  // stem_role: 0 = stand-alone, 1 = "lead" (the 7x7x7 stems carry the peer member's stem weights `peer_w` as columns 64..127
  // and store them into a persistent peer buffer), 2 = "follow" (no stem op, the first consumers read that buffer) - the
  // twin of lowering.Lowerer(stem_role=, stem_peer=)
  Lowerer(const Graph& g, const std::map<std::string, std::vector<Tensor>>& w, bool bf16, int max_batch, bool persist_input,
          const std::vector<int>& input_f32, int stem_role = 0, const std::map<std::string, std::vector<Tensor>>* peer_w = nullptr)
      : g_(g), w_(w), nb_(max_batch), persist_input_(persist_input), input_f32_(input_f32), stem_role_(stem_role), peer_w_(peer_w) {
    if (stem_role < 0 || stem_role > 2 || (stem_role == 1) != (peer_w != nullptr))
      throw std::runtime_error("stem_role is 0, 1 = lead (with the peer's weights) or 2 = follow");
    act_ = bf16 ? CSE_BF16 : CSE_F32;
    use_tc_ = bf16;
    fuse_pool_ = bf16;
    for (auto& n : g.nodes) consumers_[n.name];
    for (auto& n : g.nodes)
      for (auto& i : n.inputs) consumers_[i].push_back(n.name);
    for (auto& n : g.nodes)
      if (n.op == "concat" && n.out_shape.size() == 4) {
        int off = 0;
        for (auto& i : n.inputs) {
          if (consumers_[i].size() == 1) place_[i] = {n.name, off};
          off += g.shape(i).back();
        }
      }
  }

  Plan lower() {
    for (auto& n : g_.nodes) {
      if (done_.count(n.name)) continue;
      dispatch(n);
    }
    // liveness
    for (size_t idx = 0; idx < ops_.size(); ++idx) {
      DevOp& op = ops_[idx];
      for (TRef* r : {&op.in0, &op.in1, &op.out0, &op.out1, &op.out2})
        if (r->valid()) {
          bufs_[r->buf].first = std::min(bufs_[r->buf].first, (int)idx);
          bufs_[r->buf].last = std::max(bufs_[r->buf].last, (int)idx);
        }
      if (op.part >= 0) bufs_[op.part].first = bufs_[op.part].last = (int)idx;
    }
    if (logits_.valid()) bufs_[logits_.buf].last = (int)ops_.size() + 1;
    if (probs_.valid()) bufs_[probs_.buf].last = (int)ops_.size() + 1;
    if (persist_input_)
      for (auto& op : ops_)
        if (op.kind == CSE_OP_PREPROCESS) { bufs_[op.out0.buf].last = (int)ops_.size() + 1; bufs_[op.out0.buf].first = stem_role_ ? -2 : 0; }
    // peer stem buffers: right behind the pre-processed clips, before any plan-local buffer - the same offset in the
    // leader's and the follower's plan
    for (int b : persistent_) { bufs_[b].first = -1; bufs_[b].last = (int)ops_.size() + 1; }
    Plan plan;
    plan.workspace_bytes = assign_offsets();
    // weight arena: blobs at 256-byte aligned offsets, in creation order
    std::vector<long long> offs;
    long long cur = 0;
    for (auto& b : blobs_) { cur = round_up(cur, 256); offs.push_back(cur); cur += (long long)b.size(); }
    plan.arena.assign((size_t)round_up(std::max(cur, 256LL), 256), 0);
    for (size_t i = 0; i < blobs_.size(); ++i) std::memcpy(plan.arena.data() + offs[i], blobs_[i].data(), blobs_[i].size());
    for (auto& op : ops_) plan.ops.push_back(to_struct(op, offs));
    plan.logits_off = logits_.valid() ? byte_off(logits_) : -1;
    plan.probs_off = probs_.valid() ? byte_off(probs_) : -1;
    plan.n_inputs = (int)g_.inputs.size();
    return plan;
  }

 private:
  const Graph& g_;
  const std::map<std::string, std::vector<Tensor>>& w_;
  int nb_, act_;
  bool use_tc_, fuse_pool_, persist_input_;
  std::vector<int> input_f32_;
  int stem_role_ = 0;
  const std::map<std::string, std::vector<Tensor>>* peer_w_ = nullptr;
  std::vector<int> persistent_;
  std::vector<DevOp> ops_;
  std::vector<Buf> bufs_;
  std::vector<Blob> blobs_;
  std::map<std::string, Val> val_;
  std::map<std::string, std::vector<std::string>> consumers_;
  std::set<std::string> done_;
  std::map<std::string, std::pair<std::string, int>> place_;
  std::map<std::string, TRef> concat_ref_;
  TRef logits_, probs_;

  static int esize(int dt) { return dt == CSE_F32 ? 4 : 2; }
  long long byte_off(const TRef& r) const { return bufs_[r.buf].offset + (long long)r.coff * esize(r.dtype); }

  int new_buf(const std::string& name, const int dims[3], int ld, int dt) {
    const long long nbytes = (long long)nb_ * dims[0] * dims[1] * dims[2] * ld * esize(dt);
    Buf b; b.name = name; b.nbytes = round_up(std::max(nbytes, 16LL), ALIGN);
    bufs_.push_back(b);
    return (int)bufs_.size() - 1;
  }
  static TRef make_ref(int buf, int coff, int C, int ld, const int dims[3], int dt) {
    TRef r; r.buf = buf; r.coff = coff; r.C = C; r.ld = ld; r.dtype = dt;
    for (int i = 0; i < 3; ++i) r.dims[i] = dims[i];
    return r;
  }
  // output view for the chain ending in `final_name`: a concat slice if one was planned
  TRef out_ref(const std::string& final_name, const int dims[3], int C, int dt) {
    auto it = place_.find(final_name);
    if (it != place_.end() && dt == act_) {
      const std::string& cname = it->second.first;
      if (!concat_ref_.count(cname)) {
        const int ctot = g_.shape(cname).back();
        concat_ref_[cname] = make_ref(new_buf(cname, dims, ctot, dt), 0, ctot, ctot, dims, dt);
      }
      const TRef& base = concat_ref_[cname];
      return make_ref(base.buf, it->second.second, C, base.ld, dims, dt);
    }
    return make_ref(new_buf(final_name, dims, C, dt), 0, C, C, dims, dt);
  }
  int blob(Blob b) { blobs_.push_back(std::move(b)); return (int)blobs_.size() - 1; }
  int fblob(bool present, const std::vector<float>& a) { return present ? blob(blob_f32(a)) : -1; }

  const Node* sole_consumer(const std::string& name, const char* kind) const {
    const auto& c = consumers_.at(name);
    if (c.size() == 1 && g_.at(c[0]).op == kind) return &g_.at(c[0]);
    return nullptr;
  }
  const std::vector<Tensor>& weights(const std::string& layer) const {
    auto it = w_.find(layer);
    if (it == w_.end()) throw std::runtime_error("weights for layer " + layer + " missing");
    return it->second;
  }
  Kernel5 kernel_of(const Node& n) const {
    const Tensor& t = weights(n.name)[0];
    Kernel5 k(t.shape[0], t.shape[1], t.shape[2], t.shape[3], t.shape[4]);
    k.v = t.data;
    return k;
  }

  void dispatch(const Node& n) {
    if (n.op == "input") return lower_input(n);
    if (n.op == "conv3d") return lower_conv3d(n);
    if (n.op == "bn") return lower_bn(n);
    if (n.op == "relu") return lower_relu(n);
    if (n.op == "dropout") { val_[n.name] = val_.at(n.inputs[0]); return; }
    if (n.op == "add") return lower_add(n);
    if (n.op == "concat") return lower_concat(n);
    if (n.op == "maxpool") return lower_pool(n, CSE_OP_MAXPOOL3D, val_.at(n.inputs[0]).ref, n.pb, false, {n.name});
    if (n.op == "avgpool") { const int z[3] = {0, 0, 0}; return lower_pool(n, CSE_OP_AVGPOOL3D, val_.at(n.inputs[0]).ref, z, false, {n.name}); }
    if (n.op == "zeropad") return lower_zeropad(n);
    if (n.op == "flatten") return lower_flatten(n);
    if (n.op == "dense") return lower_dense(n);
    throw std::runtime_error("unknown layer op " + n.op);
  }

  void set_val(const std::vector<std::string>& layers, const TRef& ref) {
    for (auto& l : layers) { Val v; v.ref = ref; val_[l] = v; done_.insert(l); }
  }

  // ---- input ---------------------------------------------------------------------------------------------------
  bool pair_pool_ok(const Node& conv, int h, int w, int c) const {
    if (!(fuse_pool_ && 4 * c <= 16 && w % 2 == 0 && h % 2 == 0)) return false;
    if (conv.filters != 64 || !conv.use_bias || place_.count(conv.name)) return false;
    std::string final = conv.name;
    if (!(conv.act == ACT_NONE || conv.act == ACT_RELU)) return false;
    if (conv.act != ACT_RELU) {
      const Node* nx = sole_consumer(final, "relu");
      if (nx) final = nx->name;
    }
    if (place_.count(final)) return false;
    const Node* mp = sole_consumer(final, "maxpool");
    return mp && !mp->same && mp->k[0] == 1 && mp->k[1] == 2 && mp->k[2] == 2 && mp->s[0] == 1 && mp->s[1] == 2 && mp->s[2] == 2 &&
           !place_.count(mp->name);
  }

  void emit_pre(const Node& node, const TRef& out, int idx, int src_dt) {
    DevOp op; op.kind = CSE_OP_PREPROCESS; op.name = node.name; op.out0 = out; op.ext_input = idx; op.src_dtype = src_dt;
    for (int i = 0; i < 4; ++i) op.src_dims[i] = node.out_shape[i];
    ops_.push_back(op);
    Val v; v.ref = out; val_[node.name] = v;
  }

  void lower_input(const Node& node) {
    const int t = node.out_shape[0], h = node.out_shape[1], w = node.out_shape[2], c = node.out_shape[3];
    int idx = 0;
    for (size_t i = 0; i < g_.inputs.size(); ++i) if (g_.inputs[i] == node.name) idx = (int)i;
    const int src_dt = (idx < (int)input_f32_.size() && input_f32_[idx]) ? CSE_F32 : CSE_U8;
    const int ld = act_ == CSE_F32 ? c : 8;
    const auto& cons = consumers_.at(node.name);
    const Node* first = cons.size() == 1 ? &g_.at(cons[0]) : nullptr;
    const bool k333 = first && first->op == "conv3d" && first->k[0] == 3 && first->k[1] == 3 && first->k[2] == 3;
    const bool s111 = first && first->s[0] == 1 && first->s[1] == 1 && first->s[2] == 1;
    if (use_tc_ && src_dt == CSE_U8 && k333 && s111 && c <= 8 && first->same && first->filters % 8 == 0 && c <= 4) {
      if (pair_pool_ok(*first, h, w, c)) {
        const int dims[3] = {t, h, w / 2};
        TRef out = make_ref(new_buf(node.name, dims, 16, act_), 0, c, 16, dims, act_);
        out.unroll_w = 4;
        return emit_pre(node, out, idx, src_dt);
      }
      const int dims[3] = {t, h, w};
      TRef out = make_ref(new_buf(node.name, dims, 16, act_), 0, c, 16, dims, act_);
      out.unroll_w = 3;
      return emit_pre(node, out, idx, src_dt);
    }
    const bool k777 = first && first->op == "conv3d" && first->k[0] == 7 && first->k[1] == 7 && first->k[2] == 7;
    const bool s222 = first && first->s[0] == 2 && first->s[1] == 2 && first->s[2] == 2;
    if (use_tc_ && k777 && s222 && c <= 4 && first->same && first->filters % 8 == 0 && first->filters >= 16 &&
        (src_dt == CSE_U8 || c <= 2)) {
      const int h2 = (h + 1) / 2, w2 = (w + 1) / 2;
      const int cell = (int)round_up(4 * c, 8);
      const int wpad = (first->pb[2] + 1) / 2, wpitch = w2 + 3;
      if (c == 1 || c == 2) {                 // 2x2x2 cells (see lowering._input for the measured choice)
        const int t2 = (t + 1) / 2, cell3 = 8 * c;
        const int bd[3] = {t2, h2, wpitch}, dims[3] = {t2, h2, w2};
        TRef out = make_ref(new_buf(node.name, bd, cell3, act_), 0, cell3, cell3, dims, act_);
        out.wpitch = wpitch; out.wpad = wpad; out.s2d = 2; out.src_c = c;
        return emit_pre(node, out, idx, src_dt);
      }
      const int bd[3] = {t, h2, wpitch}, dims[3] = {t, h2, w2};
      TRef out = make_ref(new_buf(node.name, bd, cell, act_), 0, cell, cell, dims, act_);
      out.wpitch = wpitch; out.wpad = wpad; out.s2d = 1; out.src_c = c;
      return emit_pre(node, out, idx, src_dt);
    }
    const int dims[3] = {t, h, w};
    TRef out = make_ref(new_buf(node.name, dims, ld, act_), 0, c, ld, dims, act_);
    emit_pre(node, out, idx, src_dt);
  }

  // ---- convolutions ------------------------------------------------------------------------------------------------
  bool tc_ok(const TRef& x, int co, const int s[3], int out_dtype, const TRef* residual) const {
    (void)s;
    return use_tc_ && x.dtype == CSE_BF16 && out_dtype == CSE_BF16 && x.C % 8 == 0 && x.ld % 8 == 0 && x.coff % 8 == 0 &&
           co % 8 == 0 && co >= 16 && (!residual || (residual->ld % 8 == 0 && residual->coff % 8 == 0));
  }

  int split_k_factor(int ci, const Kernel5& k, const int out_dims[3]) const {
    if (!use_tc_) return 1;
    const int kc = choose_kc(ci);
    const int mult[3] = {1, 1, 1};
    int brick[4];
    choose_brick(nb_, out_dims[0], out_dims[1], out_dims[2], mult, brick);
    const int m_tiles = cdiv(nb_, brick[0]) * cdiv(out_dims[0], brick[1]) * cdiv(out_dims[1], brick[2]) * cdiv(out_dims[2], brick[3]);
    int bn, n_tiles;
    choose_bn(k.co, m_tiles, &bn, &n_tiles);
    const int ksteps = k.kd * k.kh * k.kw * cdiv(ci, kc);
    const int tiles = m_tiles * n_tiles;
    if (tiles * 2 > SM_COUNT || ksteps < 16) return 1;
    const int ks = std::min(std::min(SM_COUNT / tiles, ksteps / 8), 8);
    return ks >= 2 ? ks : 1;
  }

  struct PoolFuse { bool ok = false; int k[3], dims[3]; bool zero = false; std::vector<std::string> names; };
  PoolFuse fusable_pool(const std::string& final, const int out_dims[3]) const {
    PoolFuse pf;
    if (!fuse_pool_) return pf;
    std::string cur = final;
    const Node* zp = sole_consumer(cur, "zeropad");
    if (zp) {
      for (int i = 0; i < 3; ++i) if (zp->pads[i][0] != 0) return pf;
      pf.names.push_back(zp->name);
      pf.zero = true;
      cur = zp->name;
    }
    const Node* mp = sole_consumer(cur, "maxpool");
    if (!mp || mp->same) return pf;
    for (int i = 0; i < 3; ++i) if (mp->k[i] != mp->s[i]) return pf;
    for (int i = 0; i < 3; ++i) if (128 % mp->k[i]) return pf;
    if (mp->k[0] * mp->k[1] * mp->k[2] > 16) return pf;
    if (place_.count(mp->name) || (zp && place_.count(zp->name))) return pf;
    for (int i = 0; i < 3; ++i) { pf.k[i] = mp->k[i]; pf.dims[i] = mp->out_shape[i]; }
    pf.names.push_back(mp->name);
    pf.ok = true;
    (void)out_dims;
    return pf;
  }

  struct Chain { const std::vector<Tensor>* bn = nullptr; bool bn_gamma = false; bool relu = false; std::string final; std::vector<std::string> layers; };
  Chain conv_chain(const Node& node) const {
    Chain ch;
    ch.layers = {node.name};
    ch.final = node.name;
    ch.relu = node.act == ACT_RELU;
    const Node* nx = sole_consumer(ch.final, "bn");
    if (nx && !ch.relu) { ch.bn = &weights(nx->name); ch.bn_gamma = nx->bn_scale; ch.layers.push_back(nx->name); ch.final = nx->name; }
    nx = sole_consumer(ch.final, "relu");
    if (nx && !ch.relu) { ch.relu = true; ch.layers.push_back(nx->name); ch.final = nx->name; }
    return ch;
  }

  struct Second { bool present = false; std::vector<float> sc, sh; bool relu = false; std::string name; };

  // Emit one CONV3D op.  halo: 0 none, 1 = (kd,kh)-halo request, 2 = h-halo request
  DevOp& conv_like(const std::string& name, const TRef& x, const Kernel5& kernel, const std::vector<float>* bias, const int k[3],
                   const int s[3], const int pads[3], const int out_dims[3], const std::vector<Tensor>* chain_bn, bool bn_gamma,
                   bool relu, const std::string& final_name, int out_dtype, const TRef* residual, const Second* second, int halo,
                   const PoolFuse* pool) {
    const int co = kernel.co;
    std::vector<float> scale, shift;
    bool has_scale, has_shift;
    fold_bn(bias, chain_bn, bn_gamma, &scale, &shift, &has_scale, &has_shift);
    const bool pooled = pool && pool->ok;
    DevOp op; op.kind = CSE_OP_CONV3D; op.name = name; op.in0 = x;
    if (residual) op.in1 = *residual;
    op.out0 = out_ref(final_name, pooled ? pool->dims : out_dims, co, out_dtype);
    for (int i = 0; i < 3; ++i) { op.k[i] = k[i]; op.s[i] = s[i]; op.pad[i] = pads[i]; }
    op.relu0 = relu ? 1 : 0;
    int mult[3] = {1, 1, 1};
    if (pooled) {
      for (int i = 0; i < 3; ++i) { op.pool_k[i] = pool->k[i]; op.conv_out_dims[i] = out_dims[i]; mult[i] = pool->k[i]; }
      op.pool_zero = pool->zero ? 1 : 0;
      op.has_conv_out_dims = true;
    }
    op.scale0 = fblob(has_scale, scale);
    op.shift0 = fblob(has_shift, shift);
    const int ci = x.C;
    const bool ok = tc_ok(x, co, s, out_dtype, residual) && op.out0.ld % 8 == 0 && op.out0.coff % 8 == 0;
    if (pooled && !ok) throw std::runtime_error("fused pooling was planned for a conv that cannot use the tcgen05 engine");
    if (ok) {
      const int kc = choose_kc(ci);
      int gen_brick[4];
      choose_brick(nb_, out_dims[0], out_dims[1], out_dims[2], mult, gen_brick);
      const int m_tiles = cdiv(nb_, gen_brick[0]) * cdiv(out_dims[0], gen_brick[1]) * cdiv(out_dims[1], gen_brick[2]) *
                          cdiv(out_dims[2], gen_brick[3]);
      int bn, n_tiles;
      choose_bn(co, m_tiles, &bn, &n_tiles);
      op.engine = CSE_ENGINE_TCGEN05; op.w_dtype = CSE_BF16; op.kc = kc; op.bn = bn;
      int hw_brick[4];
      const bool have_hw = (halo == 1) && choose_brick_hw(out_dims[1], out_dims[2], mult, hw_brick);
      int ph_brick[4];
      const double ph_eff = choose_brick_pair_halo(out_dims[1], out_dims[2], kernel.kh, ph_brick);
      const long long ph_tiles = (long long)nb_ * out_dims[0] * cdiv(out_dims[1], ph_brick[2]) * cdiv(out_dims[2], ph_brick[3]);
      if (!halo && !pooled && kernel.kw > 1 && s[1] == 1 && s[2] == 1 && kc == 64 && n_tiles == 1 && bn <= 128 && bn % 16 == 0 &&
          ph_eff >= 0.8 && ph_tiles >= 2 * SM_COUNT) {
        op.halo = 3;
        std::copy(ph_brick, ph_brick + 4, op.brick);
        op.w_blob = blob(pack_tc_weights_hhalo(kernel, kc, bn, n_tiles));
      } else if (halo == 2 && kernel.kw == 1 && s[1] == 1 && s[2] == 1 && kernel.kh * bn <= 256 && (bn * kc * 2) % 1024 == 0 && !pooled) {
        op.halo = 2;
        choose_brick_hhalo(out_dims[1], out_dims[2], kernel.kh, op.brick);
        op.w_blob = blob(pack_tc_weights_hhalo(kernel, kc, bn, n_tiles));
      } else if (have_hw && kernel.kw == 1 && ci <= kc && kernel.kh * bn <= 256 && (bn * kc * 2) % 1024 == 0) {
        op.halo = 1;
        std::copy(hw_brick, hw_brick + 4, op.brick);
        op.w_blob = blob(pack_tc_weights_halo(kernel, kc, bn, n_tiles));
      } else {
        std::copy(gen_brick, gen_brick + 4, op.brick);
        op.w_blob = blob(pack_tc_weights(kernel, kc, bn, n_tiles));
        const int ks = pooled ? 1 : split_k_factor(ci, kernel, out_dims);
        if (ks >= 2) {
          op.ksplit = ks;
          Buf b; b.name = name + ":splitk"; b.nbytes = round_up((long long)ks * m_tiles * 128 * n_tiles * bn * 4, ALIGN);
          bufs_.push_back(b);
          op.part = (int)bufs_.size() - 1;
        }
      }
    } else {
      op.engine = CSE_ENGINE_DIRECT;
      if (x.dtype == CSE_BF16) {
        op.w_dtype = CSE_BF16;
        std::vector<uint16_t> wb(kernel.v.size());
        for (size_t i = 0; i < wb.size(); ++i) wb[i] = bf16_bits(kernel.v[i]);
        op.w_blob = blob(blob_u16(wb));
      } else {
        op.w_dtype = CSE_F32;
        op.w_blob = blob(blob_f32(kernel.v));
      }
    }
    if (second && second->present) {
      op.out1 = out_ref(second->name, out_dims, co, out_dtype);
      op.scale1 = fblob(true, second->sc);
      op.shift1 = fblob(true, second->sh);
      op.relu1 = second->relu ? 1 : 0;
    }
    ops_.push_back(op);
    return ops_.back();
  }

  bool fused_siblings(const Node& node) {
    if (!use_tc_) return false;
    auto is_pw = [&](const Node& n) {
      return n.op == "conv3d" && n.k[0] == 1 && n.k[1] == 1 && n.k[2] == 1 && n.s[0] == 1 && n.s[1] == 1 && n.s[2] == 1 && !done_.count(n.name);
    };
    if (!is_pw(node)) return false;
    const std::string& src = node.inputs[0];
    const TRef x = val_.at(src).ref;
    std::vector<const Node*> sibs;
    for (auto& c : consumers_.at(src)) if (is_pw(g_.at(c))) sibs.push_back(&g_.at(c));
    bool mine = false;
    for (auto* n : sibs) mine = mine || n->name == node.name;
    if (sibs.size() < 2 || !mine) return false;
    std::vector<Chain> chains;
    for (auto* n : sibs) chains.push_back(conv_chain(*n));
    std::vector<int> placed;
    for (size_t i = 0; i < chains.size(); ++i) if (place_.count(chains[i].final)) placed.push_back((int)i);
    bool same_relu = true;
    for (auto& ch : chains) same_relu = same_relu && ch.relu == chains[0].relu;
    if (placed.size() != 1 || sibs.size() > 3 || !same_relu) return false;
    std::vector<int> order = {placed[0]};
    for (int i = 0; i < (int)sibs.size(); ++i) if (i != placed[0]) order.push_back(i);
    std::vector<const Node*> s2;
    std::vector<Chain> c2;
    for (int i : order) { s2.push_back(sibs[i]); c2.push_back(chains[i]); }
    std::vector<int> cos;
    for (auto* n : s2) cos.push_back(n->filters);
    if (cos[0] % 16 || (cos.size() == 3 && cos[1] % 16)) return false;
    for (int c : cos) if (c % 8) return false;
    for (size_t i = 1; i < c2.size(); ++i) if (place_.count(c2[i].final)) return false;
    int co = 0;
    for (int c : cos) co += c;
    if (!tc_ok(x, co, S1, act_, nullptr)) return false;
    const std::vector<int>& os = node.out_shape;
    const int out_dims[3] = {os[0], os[1], os[2]};
    // NOTE: out_ref allocates; lowering.py allocates the outputs before the remaining checks, and so do we
    std::vector<TRef> outs;
    for (size_t i = 0; i < c2.size(); ++i) outs.push_back(out_ref(c2[i].final, out_dims, cos[i], act_));
    for (auto& o : outs) if (o.ld % 8 || o.coff % 8) return false;
    const int ci = x.C;
    Kernel5 kernel(1, 1, 1, ci, co);
    int col = 0;
    std::vector<float> scales, shifts;
    for (size_t i = 0; i < s2.size(); ++i) {
      const Tensor& t = weights(s2[i]->name)[0];
      for (int c = 0; c < ci; ++c)
        for (int o = 0; o < cos[i]; ++o) kernel.at(0, 0, 0, c, col + o) = t.data[(size_t)c * cos[i] + o];
      const std::vector<float>* bias = s2[i]->use_bias ? &weights(s2[i]->name)[1].data : nullptr;
      std::vector<float> sc, sh;
      bool hs, hsh;
      fold_bn(bias, c2[i].bn, c2[i].bn_gamma, &sc, &sh, &hs, &hsh);
      if (!hs) sc.assign(cos[i], 1.0f);
      if (!hsh) sh.assign(cos[i], 0.0f);
      scales.insert(scales.end(), sc.begin(), sc.end());
      shifts.insert(shifts.end(), sh.begin(), sh.end());
      col += cos[i];
    }
    DevOp op; op.kind = CSE_OP_CONV3D;
    for (size_t i = 0; i < s2.size(); ++i) op.name += (i ? "+" : "") + s2[i]->name;
    op.in0 = x;
    op.out0 = make_ref(outs[0].buf, outs[0].coff, co, outs[0].ld, out_dims, act_);
    op.out1 = outs[1];
    op.relu0 = c2[0].relu ? 1 : 0;
    op.scale0 = blob(blob_f32(scales));
    op.shift0 = blob(blob_f32(shifts));
    op.out_split = cos[0];
    if (cos.size() == 3) { op.out_split2 = cos[0] + cos[1]; op.out2 = outs[2]; }
    const int kc = choose_kc(x.C);
    const int mult[3] = {1, 1, 1};
    choose_brick(nb_, out_dims[0], out_dims[1], out_dims[2], mult, op.brick);
    const int m_tiles = cdiv(nb_, op.brick[0]) * cdiv(out_dims[0], op.brick[1]) * cdiv(out_dims[1], op.brick[2]) * cdiv(out_dims[2], op.brick[3]);
    int bn, n_tiles;
    choose_bn(co, m_tiles, &bn, &n_tiles);
    for (int gran : {64, 32}) {
      const int wide = (int)round_up(bn, gran);
      if (op.out_split % gran == 0 && op.out_split2 % gran == 0 && wide <= 256 && wide * n_tiles <= 1.15 * co) { bn = wide; break; }
    }
    op.engine = CSE_ENGINE_TCGEN05; op.w_dtype = CSE_BF16; op.kc = kc; op.bn = bn;
    op.w_blob = blob(pack_tc_weights(kernel, kc, bn, n_tiles));
    ops_.push_back(op);
    for (size_t i = 0; i < c2.size(); ++i) set_val(c2[i].layers, outs[i]);
    return true;
  }

  void lower_conv3d(const Node& node) {
    if (fused_siblings(node)) return;
    const TRef x = val_.at(node.inputs[0]).ref;
    const Kernel5 kernel = kernel_of(node);
    const std::vector<float>* bias = node.use_bias ? &weights(node.name)[1].data : nullptr;
    Chain ch = conv_chain(node);
    const int out_dims[3] = {node.out_shape[0], node.out_shape[1], node.out_shape[2]};
    if (x.s2d) return s2d_stem_conv(node, x, kernel, bias, ch, out_dims);
    if (x.unroll_w == 4) return pair_pool_stem(node, x, kernel, bias, ch, out_dims);
    PoolFuse pool;
    std::string final = ch.final;
    std::vector<std::string> layers = ch.layers;
    if (x.wpitch || x.unroll_w || tc_ok(x, kernel.co, node.s, act_, nullptr)) {
      PoolFuse fp = fusable_pool(final, out_dims);
      if (fp.ok && !(x.wpitch || x.unroll_w) && split_k_factor(x.C, kernel, out_dims) > 1) fp.ok = false;
      if (fp.ok && !place_.count(final)) {
        pool = fp;
        layers.insert(layers.end(), fp.names.begin(), fp.names.end());
        final = fp.names.back();
      }
    }
    if (x.wpitch || x.unroll_w) return packed_stem_conv(node, x, kernel, bias, ch, final, layers, out_dims, pool);
    // residual fusion: conv (no bn/relu tail) whose only consumer is add([shortcut, this])
    const Node* add = (!ch.bn && !ch.relu) ? sole_consumer(final, "add") : nullptr;
    if (add && add->inputs.size() == 2 && add->inputs[1] == final && !val_.count(add->inputs[0])) {
      const Node& sc = g_.at(add->inputs[0]);      // projection shortcut: created after the residual convs, lower it first
      if (sc.op == "conv3d" && val_.count(sc.inputs[0])) lower_conv3d(sc);
    }
    if (add && add->inputs.size() == 2 && add->inputs[1] == final && val_.count(add->inputs[0]))
      return fused_residual(node, *add, x, kernel, bias, out_dims);
    DevOp& op = conv_like(node.name, x, kernel, bias, node.k, node.s, node.pb, out_dims, ch.bn, ch.bn_gamma, ch.relu, final, act_, nullptr,
                          nullptr, 0, &pool);
    set_val(layers, op.out0);
  }

  void packed_stem_conv(const Node& node, const TRef& x, const Kernel5& kernel, const std::vector<float>* bias, const Chain& ch,
                        const std::string& final, const std::vector<std::string>& layers, const int out_dims[3], const PoolFuse& pool) {
    // 3x3x3 'same' stem on a C <= 4 clip with the W-unrolled input: kw taps folded into the channel axis
    const int ci = kernel.ci, co = kernel.co;
    Kernel5 k2(3, 3, 1, 16, co);
    for (int fd = 0; fd < 3; ++fd)
      for (int fh = 0; fh < 3; ++fh)
        for (int j = 0; j < 3; ++j)
          for (int c = 0; c < ci; ++c)
            for (int o = 0; o < co; ++o) k2.at(fd, fh, 0, j * ci + c, o) = kernel.at(fd, fh, j, c, o);
    TRef view = make_ref(x.buf, 0, 16, 16, x.dims, x.dtype);
    const int k[3] = {3, 3, 1}, s[3] = {1, 1, 1}, pads[3] = {1, 1, 0};
    DevOp& op = conv_like(node.name, view, k2, bias, k, s, pads, out_dims, ch.bn, ch.bn_gamma, ch.relu, final, act_, nullptr, nullptr, 1, &pool);
    if (op.engine != CSE_ENGINE_TCGEN05) throw std::runtime_error("packed stem must lower to the tcgen05 engine");
    set_val(layers, op.out0);
  }

  void pair_pool_stem(const Node& node, const TRef& x, const Kernel5& kernel, const std::vector<float>* bias, const Chain& ch,
                      const int out_dims[3]) {
    (void)out_dims;
    const int ci = kernel.ci, co = kernel.co;
    const Node* mp = sole_consumer(ch.final, "maxpool");
    const int t = x.dims[0], h = x.dims[1], wp = x.dims[2];
    Kernel5 k2(3, 3, 1, 16, 2 * co);
    for (int px = 0; px < 2; ++px)
      for (int fw = 0; fw < 3; ++fw)
        for (int fd = 0; fd < 3; ++fd)
          for (int fh = 0; fh < 3; ++fh)
            for (int c = 0; c < ci; ++c)
              for (int o = 0; o < co; ++o) k2.at(fd, fh, 0, (px + fw) * ci + c, px * co + o) = kernel.at(fd, fh, fw, c, o);
    const int pooled[3] = {mp->out_shape[0], mp->out_shape[1], mp->out_shape[2]};
    TRef out0 = out_ref(mp->name, pooled, co, act_);
    if (out0.ld % 8 || out0.coff % 8) throw std::runtime_error("pair-pool stem output must be 16-byte aligned");
    DevOp op; op.kind = CSE_OP_CONV3D; op.name = node.name;
    op.in0 = make_ref(x.buf, 0, 16, 16, x.dims, x.dtype);
    op.out0 = out0;
    op.k[0] = 3; op.k[1] = 3; op.k[2] = 1; op.pad[0] = 1; op.pad[1] = 1; op.pad[2] = 0;
    op.relu0 = ch.relu ? 1 : 0;
    op.engine = CSE_ENGINE_TCGEN05; op.w_dtype = CSE_BF16; op.kc = 16; op.bn = 2 * co; op.halo = 1; op.pair_pool = 1;
    op.pool_k[0] = 1; op.pool_k[1] = 2; op.pool_k[2] = 1; op.pool_zero = 0;
    op.has_conv_out_dims = true; op.conv_out_dims[0] = t; op.conv_out_dims[1] = h; op.conv_out_dims[2] = wp;
    long long best = -1;
    for (int bw : {8, 16}) {
      const int bh = 128 / bw;
      const long long tiles = (long long)cdiv(h, bh) * cdiv(wp, bw);
      if (best < 0 || tiles < best) { best = tiles; op.brick[0] = 1; op.brick[1] = 1; op.brick[2] = bh; op.brick[3] = bw; }
    }
    op.w_blob = blob(pack_tc_weights_halo(k2, 16, 2 * co, 1));
    op.shift0 = blob(blob_f32(*bias));
    ops_.push_back(op);
    std::vector<std::string> layers = ch.layers;
    layers.push_back(mp->name);
    set_val(layers, out0);
  }

  void s2d_stem_conv(const Node& node, const TRef& x, const Kernel5& kernel, const std::vector<float>* bias, const Chain& ch,
                     const int out_dims[3]) {
    const int ci = kernel.ci, co = kernel.co, cell = x.ld;
    const int* pb = node.pb;
    int off[3];
    for (int i = 0; i < 3; ++i) off[i] = pb[i] - 2 * ((pb[i] + 1) / 2);
    const bool depth = x.s2d == 2;
    auto regroup = [&](const Kernel5& kern) {
      Kernel5 k2(depth ? 4 : 7, 4, 1, 4 * cell, co);
      for (int fd = 0; fd < k2.kd; ++fd)
        for (int pd = 0; pd < (depth ? 2 : 1); ++pd) {
          const int td = depth ? 2 * fd + pd + off[0] : fd;
          if (td < 0 || td >= 7) continue;
          for (int fh = 0; fh < 4; ++fh)
            for (int ph = 0; ph < 2; ++ph) {
              const int th = 2 * fh + ph + off[1];
              if (th < 0 || th >= 7) continue;
              for (int fw = 0; fw < 4; ++fw)
                for (int pw = 0; pw < 2; ++pw) {
                  const int tw = 2 * fw + pw + off[2];
                  if (tw < 0 || tw >= 7) continue;
                  const int c0 = fw * cell + ((pd * 2 + ph) * 2 + pw) * ci;
                  for (int c = 0; c < ci; ++c)
                    for (int o = 0; o < co; ++o) k2.at(fd, fh, 0, c0 + c, o) = kern.at(td, th, tw, c, o);
                }
            }
        }
      return k2;
    };
    TRef view = make_ref(x.buf, 0, 4 * cell, cell, x.dims, x.dtype);
    view.wpitch = x.wpitch; view.wpad = x.wpad;
    int k[3], s[3], pads[3];
    if (depth) { k[0] = 4; k[1] = 4; k[2] = 1; s[0] = s[1] = s[2] = 1; pads[0] = (pb[0] + 1) / 2; pads[1] = (pb[1] + 1) / 2; pads[2] = 0; }
    else { k[0] = 7; k[1] = 4; k[2] = 1; s[0] = 2; s[1] = 1; s[2] = 1; pads[0] = pb[0]; pads[1] = (pb[1] + 1) / 2; pads[2] = 0; }
    if (stem_role_ != 0) {
      if (!use_tc_ || co != 64) throw std::runtime_error("stem_role needs the bf16 tcgen05 path and a 64-filter 7x7x7 stem");
      const int pbuf = new_buf(ch.final + ":peer", out_dims, co, act_);
      persistent_.push_back(pbuf);
      const TRef peer = make_ref(pbuf, 0, co, co, out_dims, act_);
      if (stem_role_ == 2) { set_val(ch.layers, peer); return; }
      auto pit = peer_w_->find(node.name);
      if (pit == peer_w_->end()) throw std::runtime_error("peer weights for layer " + node.name + " missing");
      const Tensor& pt = pit->second[0];
      Kernel5 pkernel(pt.shape[0], pt.shape[1], pt.shape[2], pt.shape[3], pt.shape[4]);
      pkernel.v = pt.data;
      const std::vector<float>* pbias = node.use_bias ? &pit->second[1].data : nullptr;
      const std::vector<Tensor>* pbn = nullptr;
      if (ch.bn) {
        auto bit = peer_w_->find(ch.layers[1]);
        if (bit == peer_w_->end()) throw std::runtime_error("peer weights for layer " + ch.layers[1] + " missing");
        pbn = &bit->second;
      }
      std::vector<float> scale, shift;
      for (int who = 0; who < 2; ++who) {
        std::vector<float> sc, sh;
        bool hs, hsh;
        fold_bn(who ? pbias : bias, who ? pbn : ch.bn, ch.bn_gamma, &sc, &sh, &hs, &hsh);
        if (!hs) sc.assign(co, 1.f);
        if (!hsh) sh.assign(co, 0.f);
        scale.insert(scale.end(), sc.begin(), sc.end());
        shift.insert(shift.end(), sh.begin(), sh.end());
      }
      const Kernel5 ka = regroup(kernel), kb = regroup(pkernel);
      Kernel5 kcat(ka.kd, ka.kh, ka.kw, ka.ci, 2 * co);
      for (int a = 0; a < ka.kd; ++a)
        for (int b = 0; b < ka.kh; ++b)
          for (int c = 0; c < ka.ci; ++c)
            for (int o = 0; o < co; ++o) {
              kcat.at(a, b, 0, c, o) = ka.at(a, b, 0, c, o);
              kcat.at(a, b, 0, c, co + o) = kb.at(a, b, 0, c, o);
            }
      const TRef o0 = out_ref(ch.final, out_dims, co, act_);
      DevOp op; op.kind = CSE_OP_CONV3D; op.name = node.name + "+peer"; op.in0 = view;
      op.out0 = make_ref(o0.buf, o0.coff, 2 * co, o0.ld, out_dims, act_);
      op.out1 = peer;
      for (int i = 0; i < 3; ++i) { op.k[i] = k[i]; op.s[i] = s[i]; op.pad[i] = pads[i]; }
      op.relu0 = ch.relu ? 1 : 0;
      op.scale0 = blob(blob_f32(scale));
      op.shift0 = blob(blob_f32(shift));
      op.out_split = co;
      op.engine = CSE_ENGINE_TCGEN05; op.w_dtype = CSE_BF16; op.kc = choose_kc(view.C); op.bn = 2 * co; op.halo = 3;
      choose_brick_hhalo(out_dims[1], out_dims[2], 4, op.brick);
      op.w_blob = blob(pack_tc_weights_hhalo(kcat, op.kc, op.bn, 1));
      ops_.push_back(op);
      set_val(ch.layers, o0);
      return;
    }
    DevOp& op = conv_like(node.name, view, regroup(kernel), bias, k, s, pads, out_dims, ch.bn, ch.bn_gamma, ch.relu, ch.final, act_, nullptr, nullptr, 2, nullptr);
    if (op.engine != CSE_ENGINE_TCGEN05) throw std::runtime_error("s2d stem must lower to the tcgen05 engine");
    set_val(ch.layers, op.out0);
  }

  void fused_residual(const Node& node, const Node& add, const TRef& x, const Kernel5& kernel, const std::vector<float>* bias,
                      const int out_dims[3]) {
    const TRef sc = val_.at(add.inputs[0]).ref;
    std::vector<std::string> layers = {node.name, add.name}, second_layers;
    Second second;
    std::vector<const Node*> bn_nodes;
    for (auto& c : consumers_.at(add.name)) if (g_.at(c).op == "bn") bn_nodes.push_back(&g_.at(c));
    if (bn_nodes.size() == 1) {
      const Node* bn = bn_nodes[0];
      const Node* r = sole_consumer(bn->name, "relu");
      bool hs, hsh;
      fold_bn(nullptr, &weights(bn->name), bn->bn_scale, &second.sc, &second.sh, &hs, &hsh);
      second.present = true;
      second.relu = r != nullptr;
      second.name = r ? r->name : bn->name;
      second_layers.push_back(bn->name);
      if (r) second_layers.push_back(r->name);
    }
    DevOp& op = conv_like(node.name, x, kernel, bias, node.k, node.s, node.pb, out_dims, nullptr, false, false, add.name, act_, &sc,
                          &second, 0, nullptr);
    const TRef o0 = op.out0, o1 = op.out1;
    set_val(layers, o0);
    set_val(second_layers, o1);
  }

  // ---- the other layers ----------------------------------------------------------------------------------------------
  void lower_bn(const Node& node) {
    const TRef x = val_.at(node.inputs[0]).ref;
    std::vector<float> sc, sh;
    bool hs, hsh;
    fold_bn(nullptr, &weights(node.name), node.bn_scale, &sc, &sh, &hs, &hsh);
    std::vector<std::string> layers = {node.name};
    std::string final = node.name;
    bool relu = false;
    const Node* nx = sole_consumer(final, "relu");
    if (nx) { relu = true; final = nx->name; layers.push_back(final); }
    DevOp op; op.kind = CSE_OP_AFFINE; op.name = node.name; op.in0 = x;
    op.out0 = out_ref(final, x.dims, x.C, x.dtype);
    op.relu0 = relu ? 1 : 0;
    op.scale0 = fblob(hs, sc);
    op.shift0 = fblob(hsh, sh);
    ops_.push_back(op);
    set_val(layers, op.out0);
  }
  void lower_relu(const Node& node) {
    const TRef x = val_.at(node.inputs[0]).ref;
    DevOp op; op.kind = CSE_OP_AFFINE; op.name = node.name; op.in0 = x;
    op.out0 = out_ref(node.name, x.dims, x.C, x.dtype);
    op.relu0 = 1;
    ops_.push_back(op);
    Val v; v.ref = op.out0; val_[node.name] = v;
  }
  void lower_add(const Node& node) {
    const TRef a = val_.at(node.inputs[0]).ref, b = val_.at(node.inputs[1]).ref;
    DevOp op; op.kind = CSE_OP_ADD; op.name = node.name; op.in0 = a; op.in1 = b;
    op.out0 = out_ref(node.name, a.dims, a.C, a.dtype);
    ops_.push_back(op);
    Val v; v.ref = op.out0; val_[node.name] = v;
  }
  void lower_concat(const Node& node) {
    if (node.out_shape.size() == 1) {
      Val v; v.is_list = true;
      for (auto& i : node.inputs) v.list.push_back(val_.at(i).ref);
      val_[node.name] = v;
      return;
    }
    if (!concat_ref_.count(node.name)) throw std::runtime_error("concat " + node.name + ": an input was not written in place");
    for (auto& i : node.inputs) if (!place_.count(i)) throw std::runtime_error("concat " + node.name + ": an input was not written in place");
    Val v; v.ref = concat_ref_[node.name]; val_[node.name] = v;
  }
  void lower_pool(const Node& node, int kind, const TRef& x, const int pads[3], bool pad_is_zero, const std::vector<std::string>& layers) {
    const int out_dims[3] = {node.out_shape[0], node.out_shape[1], node.out_shape[2]};
    DevOp op; op.kind = kind; op.name = node.name; op.in0 = x;
    op.out0 = out_ref(node.name, out_dims, x.C, x.dtype);
    for (int i = 0; i < 3; ++i) { op.k[i] = node.k[i]; op.s[i] = node.s[i]; op.pad[i] = pads[i]; }
    op.pad_is_zero = pad_is_zero ? 1 : 0;
    ops_.push_back(op);
    set_val(layers, op.out0);
  }
  void lower_zeropad(const Node& node) {
    const Node* nx = sole_consumer(node.name, "maxpool");
    if (!nx || nx->same) throw std::runtime_error("ZeroPadding3D is only supported in front of a 'valid' MaxPooling3D");
    const int pads[3] = {node.pads[0][0], node.pads[1][0], node.pads[2][0]};
    lower_pool(*nx, CSE_OP_MAXPOOL3D, val_.at(node.inputs[0]).ref, pads, true, {node.name, nx->name});
  }
  void lower_flatten(const Node& node) {
    const TRef x = val_.at(node.inputs[0]).ref;
    if (x.ld != x.C || x.coff != 0) throw std::runtime_error("flatten of a channel slice");
    const int n = x.dims[0] * x.dims[1] * x.dims[2] * x.C;
    const int one[3] = {1, 1, 1};
    Val v; v.ref = make_ref(x.buf, 0, n, n, one, x.dtype);
    val_[node.name] = v;
  }
  void lower_dense(const Node& node) {
    const std::vector<Tensor>& wt = weights(node.name);
    const Tensor& kernel = wt[0];
    const std::vector<float>& bias = wt[1].data;
    const bool is_final = node.name == g_.output;
    const int units = node.units;
    const Val& xin = val_.at(node.inputs[0]);
    std::vector<TRef> parts = xin.is_list ? xin.list : std::vector<TRef>{xin.ref};
    const int out_dtype = is_final ? CSE_F32 : act_;
    int row = 0;
    TRef prev;
    const int one[3] = {1, 1, 1}, zero3[3] = {0, 0, 0};
    for (size_t pi = 0; pi < parts.size(); ++pi) {
      const TRef& x = parts[pi];
      Kernel5 kp(1, 1, 1, x.C, units);
      std::copy(kernel.data.begin() + (size_t)row * units, kernel.data.begin() + (size_t)(row + x.C) * units, kp.v.begin());
      row += x.C;
      const bool last = pi + 1 == parts.size();
      const std::string name = parts.size() == 1 ? node.name : node.name + "#" + std::to_string(pi);
      const std::string final_name = last ? node.name : name;
      DevOp& op = conv_like(name, x, kp, pi == 0 ? &bias : nullptr, K1, S1, zero3, one, nullptr, false, node.act == ACT_RELU && last,
                            final_name, out_dtype, prev.valid() ? &prev : nullptr, nullptr, 0, nullptr);
      prev = op.out0;
    }
    Val v; v.ref = prev; val_[node.name] = v;
    if (is_final) {
      if (node.act != ACT_SOFTMAX) throw std::runtime_error("final activation must be softmax");
      const int pb = new_buf(node.name + ":probs", one, units, CSE_F32);
      TRef probs = make_ref(pb, 0, units, units, one, CSE_F32);
      DevOp op; op.kind = CSE_OP_SOFTMAX; op.name = node.name + ":softmax"; op.in0 = prev; op.out0 = probs; op.softmax_C = units;
      ops_.push_back(op);
      logits_ = prev; probs_ = probs;
    }
  }

  // ---- buffer planning ---------------------------------------------------------------------------------------------------
  long long assign_offsets() {
    std::vector<int> live;
    for (size_t i = 0; i < bufs_.size(); ++i) if (bufs_[i].last >= 0) live.push_back((int)i);
    std::stable_sort(live.begin(), live.end(), [&](int a, int b) {
      if (bufs_[a].first != bufs_[b].first) return bufs_[a].first < bufs_[b].first;
      return bufs_[a].nbytes > bufs_[b].nbytes;
    });
    std::vector<int> placed;
    long long top = 0;
    for (int bi : live) {
      Buf& b = bufs_[bi];
      std::vector<std::pair<long long, long long>> busy;
      for (int pi : placed) {
        const Buf& p = bufs_[pi];
        if (!(p.last < b.first || p.first > b.last)) busy.push_back({p.offset, p.offset + p.nbytes});
      }
      std::sort(busy.begin(), busy.end());
      long long off = 0;
      for (auto& iv : busy) {
        if (off + b.nbytes <= iv.first) break;
        off = std::max(off, iv.second);
      }
      b.offset = off;
      placed.push_back(bi);
      top = std::max(top, off + b.nbytes);
    }
    return round_up(top, ALIGN);
  }

  cse_op to_struct(const DevOp& op, const std::vector<long long>& bo) const {
    cse_op s;
    std::memset(&s, 0, sizeof(s));
    s.kind = op.kind; s.engine = op.engine; s.w_dtype = op.w_dtype;
    if (op.kind == CSE_OP_SOFTMAX) {
      s.in_dims[0] = s.in_dims[1] = s.in_dims[2] = 1; s.in_dims[3] = op.softmax_C;
      s.out_dims[0] = s.out_dims[1] = s.out_dims[2] = 1; s.out_dims[3] = op.softmax_C;
      s.in_dtype = s.out_dtype = CSE_F32;
      s.in0_off = byte_off(op.in0); s.out0_off = byte_off(op.out0);
      s.in1_off = s.out1_off = s.w_off = s.part_off = -1;
      s.scale0_off = s.shift0_off = s.scale1_off = s.shift1_off = -1;
      return s;
    }
    if (op.in0.valid()) {
      for (int i = 0; i < 3; ++i) s.in_dims[i] = op.in0.dims[i];
      s.in_dims[3] = op.in0.C;
      s.in_ld = op.in0.ld; s.in_dtype = op.in0.dtype; s.in0_off = byte_off(op.in0); s.in_wpitch = op.in0.wpitch;
    } else {
      s.in0_off = -1;
    }
    for (int i = 0; i < 3; ++i) s.out_dims[i] = op.has_conv_out_dims ? op.conv_out_dims[i] : op.out0.dims[i];
    s.out_dims[3] = op.out0.C * (op.pair_pool ? 2 : 1);
    s.tc_pair_pool = op.pair_pool;
    s.out_split = op.out_split; s.out_split2 = op.out_split2;
    s.out_ld = op.out0.ld; s.out_dtype = op.out0.dtype; s.out0_off = byte_off(op.out0);
    if (op.pool_k[0] > 0) {
      for (int i = 0; i < 3; ++i) { s.pool_k[i] = op.pool_k[i]; s.pool_dims[i] = op.out0.dims[i]; }
      s.pool_zero = op.pool_zero;
    }
    if (op.kind == CSE_OP_PREPROCESS) {
      s.out_wpitch = op.out0.wpitch; s.out_wpad = op.out0.wpad; s.pre_unroll_w = op.out0.unroll_w; s.pre_s2d = op.out0.s2d;
      s.in_dtype = op.src_dtype;
    }
    if (op.in1.valid()) { s.in1_ld = op.in1.ld; s.in1_off = byte_off(op.in1); } else s.in1_off = -1;
    if (op.out1.valid()) { s.out1_ld = op.out1.ld; s.out1_off = byte_off(op.out1); } else s.out1_off = -1;
    if (op.out2.valid()) { s.out2_ld = op.out2.ld; s.out2_off = byte_off(op.out2); } else s.out2_off = -1;
    for (int i = 0; i < 3; ++i) { s.k[i] = op.k[i]; s.s[i] = op.s[i]; s.pad[i] = op.pad[i]; }
    s.relu0 = op.relu0; s.relu1 = op.relu1; s.pad_is_zero = op.pad_is_zero; s.ext_input = op.ext_input;
    for (int i = 0; i < 4; ++i) { s.src_dims[i] = op.src_dims[i]; s.pre_mean[i] = 0.f; s.pre_scale[i] = 1.f; }
    s.kc = op.kc; s.bn = op.bn; s.tc_halo = op.halo;
    for (int i = 0; i < 4; ++i) s.brick[i] = op.brick[i];
    s.w_off = op.w_blob >= 0 ? bo[op.w_blob] : -1;
    s.scale0_off = op.scale0 >= 0 ? bo[op.scale0] : -1;
    s.shift0_off = op.shift0 >= 0 ? bo[op.shift0] : -1;
    s.scale1_off = op.scale1 >= 0 ? bo[op.scale1] : -1;
    s.shift1_off = op.shift1 >= 0 ? bo[op.shift1] : -1;
    s.ksplit = op.ksplit;
    s.part_off = op.part >= 0 ? bufs_[op.part].offset : -1;
    s.part_bytes = op.part >= 0 ? bufs_[op.part].nbytes : 0;
    return s;
  }
};

}  // namespace mdl
}  // namespace cse

// =============================================================================================================== C ABI
struct cse_model {
  cse::mdl::Graph graph;
  std::vector<std::string> layers;                                   // weighted layers, Keras model.layers order
  std::map<std::string, std::vector<cse::mdl::Tensor>> weights;
  bool bf16 = true, persist_input = false, lowered = false, finalized = false;
  int stem_role = 0;                                                  // 0 stand-alone, 1 lead, 2 follow (cse_model_pair_stems)
  std::map<std::string, std::vector<cse::mdl::Tensor>> peer_weights;  // lead: the follower's Keras-layout tensors
  std::string model_type;
  int T = 0, H = 0, W = 0;
  int max_batch = 0, nb_classes = 0;
  std::vector<int> input_f32;
  cse::mdl::Plan plan;
  cse_plan* exec = nullptr;
  void* d_weights = nullptr;
  void* d_workspace = nullptr;
  bool own_workspace = false;
};

using namespace cse;

static int guarded(const std::function<int()>& fn) {
  try {
    return fn();
  } catch (const std::exception& e) {
    set_error("model: %s", e.what());
    return CSE_ERR_INVALID;
  }
}

extern "C" {

int cse_model_create(cse_model** out, const char* model_type, int T, int H, int W, int nb_classes, int dtype, int max_batch) {
  CSE_REQUIRE(out && model_type, "model_create: NULL argument");
  CSE_REQUIRE(dtype == CSE_BF16 || dtype == CSE_F32, "model_create: dtype must be CSE_BF16 (tensor-core path) or CSE_F32");
  CSE_REQUIRE(T >= 1 && H >= 1 && W >= 1 && nb_classes >= 1 && max_batch >= 1, "model_create: bad shape / batch");
  const std::string mt = model_type;
  bool basic;
  std::vector<int> reps;
  CSE_REQUIRE(mt == "C3D" || mt == "I3D" || mt == "TWOSTREAM_I3D" || mdl::r3d_repetitions(mt, &basic, &reps), "Unknown model %s", model_type);
  return guarded([&]() {
    std::unique_ptr<cse_model> m(new cse_model());
    m->graph = mdl::build_model_graph(mt, {T, H, W, mt == "TWOSTREAM_I3D" ? 0 : 3}, nb_classes);
    m->layers = m->graph.weighted_layers();
    for (auto& l : m->layers) {
      const mdl::Node& n = m->graph.at(l);
      std::vector<mdl::Tensor> ts(n.weights.size());
      for (size_t i = 0; i < ts.size(); ++i) ts[i].shape = n.weights[i].shape;
      m->weights[l] = std::move(ts);
    }
    m->bf16 = dtype == CSE_BF16;
    m->model_type = mt; m->T = T; m->H = H; m->W = W;
    m->max_batch = max_batch;
    m->nb_classes = nb_classes;
    m->input_f32.assign(m->graph.inputs.size(), 0);
    *out = m.release();
    return (int)CSE_OK;
  });
}

int cse_model_set_option(cse_model* m, const char* key, int value) {
  CSE_REQUIRE(m && key, "model_set_option: NULL argument");
  if (m->lowered) { set_error("model_set_option: model already lowered"); return CSE_ERR_STATE; }
  const std::string k = key;
  if (k == "persist_input") { m->persist_input = value != 0; return CSE_OK; }
  if (k == "flow_input_f32") {
    CSE_REQUIRE(m->input_f32.size() == 2, "model_set_option: flow_input_f32 needs a two-stream model");
    m->input_f32[1] = value != 0;
    return CSE_OK;
  }
  set_error("model_set_option: unknown key '%s'", key);
  return CSE_ERR_INVALID;
}

int cse_model_num_layers(const cse_model* m) { return m ? (int)m->layers.size() : 0; }

int cse_model_layer_info(const cse_model* m, int layer, char* name, int name_cap, int* n_tensors) {
  CSE_REQUIRE(m && layer >= 0 && layer < (int)m->layers.size(), "model_layer_info: layer %d out of range", layer);
  const std::string& l = m->layers[layer];
  if (name && name_cap > 0) { strncpy(name, l.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (n_tensors) *n_tensors = (int)m->graph.at(l).weights.size();
  return CSE_OK;
}

int cse_model_tensor_info(const cse_model* m, int layer, int tensor, int64_t* dims, int* ndim, char* name, int name_cap) {
  CSE_REQUIRE(m && layer >= 0 && layer < (int)m->layers.size(), "model_tensor_info: layer %d out of range", layer);
  const mdl::Node& n = m->graph.at(m->layers[layer]);
  CSE_REQUIRE(tensor >= 0 && tensor < (int)n.weights.size(), "model_tensor_info: tensor %d out of range", tensor);
  const mdl::WeightSpec& w = n.weights[tensor];
  if (ndim) *ndim = (int)w.shape.size();
  if (dims) for (size_t i = 0; i < w.shape.size(); ++i) dims[i] = w.shape[i];
  if (name && name_cap > 0) { strncpy(name, w.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  return CSE_OK;
}

int cse_model_set_weight(cse_model* m, int layer, int tensor, const float* host, const int64_t* dims, int ndim) {
  CSE_REQUIRE(m && host && dims, "model_set_weight: NULL argument");
  if (m->lowered) { set_error("model_set_weight: model already lowered"); return CSE_ERR_STATE; }
  CSE_REQUIRE(layer >= 0 && layer < (int)m->layers.size(), "model_set_weight: layer %d out of range", layer);
  std::vector<mdl::Tensor>& ts = m->weights[m->layers[layer]];
  CSE_REQUIRE(tensor >= 0 && tensor < (int)ts.size(), "model_set_weight: layer %s has %zu tensors, got index %d",
              m->layers[layer].c_str(), ts.size(), tensor);
  mdl::Tensor& t = ts[tensor];
  bool same = ndim == (int)t.shape.size();
  size_t n = 1;
  for (int i = 0; same && i < ndim; ++i) { same = dims[i] == t.shape[i]; n *= (size_t)t.shape[i]; }
  CSE_REQUIRE(same, "model_set_weight: shape mismatch for tensor %d of layer %s (positional load_weights rule)", tensor,
              m->layers[layer].c_str());
  t.data.assign(host, host + n);
  t.set = true;
  return CSE_OK;
}

int cse_model_pair_stems(cse_model* lead, cse_model* follow) {
  CSE_REQUIRE(lead && follow && lead != follow, "model_pair_stems: two distinct models");
  CSE_REQUIRE(!lead->lowered && !follow->lowered, "model_pair_stems: pair the members before lowering them");
  CSE_REQUIRE(lead->stem_role == 0 && follow->stem_role == 0, "model_pair_stems: a member is already paired");
  CSE_REQUIRE(lead->model_type == follow->model_type && lead->T == follow->T && lead->H == follow->H && lead->W == follow->W &&
                  lead->max_batch == follow->max_batch && lead->nb_classes == follow->nb_classes && lead->bf16 && follow->bf16 &&
                  lead->input_f32 == follow->input_f32,
              "model_pair_stems: both members must be the same bf16 architecture, clip shape, batch and input types");
  // every clip input must feed a 64-filter 7x7x7 / stride-2 conv (I3D, TwoStream-I3D, R3D), cf. Lowerer.stem_fusable
  for (auto& in : lead->graph.inputs) {
    const mdl::Node* conv = nullptr;
    int n_cons = 0;
    for (auto& n : lead->graph.nodes)
      for (auto& i : n.inputs)
        if (i == in) { conv = &n; ++n_cons; }
    CSE_REQUIRE(n_cons == 1 && conv->op == "conv3d" && conv->k[0] == 7 && conv->k[1] == 7 && conv->k[2] == 7 && conv->s[0] == 2 &&
                    conv->s[1] == 2 && conv->s[2] == 2 && conv->filters == 64,
                "model_pair_stems: %s has no 64-filter 7x7x7 / stride-2 stem", lead->model_type.c_str());
  }
  for (auto& l : follow->layers)
    for (size_t i = 0; i < follow->weights[l].size(); ++i)
      CSE_REQUIRE(follow->weights[l][i].set, "model_pair_stems: set the follower's weights first (tensor %zu of layer %s)", i, l.c_str());
  lead->peer_weights = follow->weights;
  lead->stem_role = 1;
  follow->stem_role = 2;
  lead->persist_input = follow->persist_input = true;
  return CSE_OK;
}

int cse_model_lower(cse_model* m) {
  CSE_REQUIRE(m, "model_lower: NULL model");
  if (m->lowered) return CSE_OK;
  for (auto& l : m->layers)
    for (size_t i = 0; i < m->weights[l].size(); ++i)
      CSE_REQUIRE(m->weights[l][i].set, "model_lower: tensor %zu of layer %s was never set", i, l.c_str());
  return guarded([&]() {
    mdl::Lowerer low(m->graph, m->weights, m->bf16, m->max_batch, m->persist_input, m->input_f32, m->stem_role,
                     m->stem_role == 1 ? &m->peer_weights : nullptr);
    m->plan = low.lower();
    m->lowered = true;
    m->weights.clear();             // the packed arena replaces the Keras-layout copies
    m->peer_weights.clear();
    return (int)CSE_OK;
  });
}

int cse_model_num_ops(const cse_model* m) { return (m && m->lowered) ? (int)m->plan.ops.size() : 0; }
int cse_model_get_op(const cse_model* m, int i, cse_op* out) {
  CSE_REQUIRE(m && m->lowered && out && i >= 0 && i < (int)m->plan.ops.size(), "model_get_op: bad argument");
  *out = m->plan.ops[i];
  return CSE_OK;
}
size_t cse_model_workspace_bytes(const cse_model* m) { return (m && m->lowered) ? (size_t)m->plan.workspace_bytes : 0; }
size_t cse_model_weight_bytes(const cse_model* m) { return (m && m->lowered) ? m->plan.arena.size() : 0; }
int cse_model_copy_weight_arena(const cse_model* m, void* host_dst, size_t cap) {
  CSE_REQUIRE(m && m->lowered && host_dst && cap >= m->plan.arena.size(), "model_copy_weight_arena: bad argument");
  memcpy(host_dst, m->plan.arena.data(), m->plan.arena.size());
  return CSE_OK;
}
int64_t cse_model_logits_offset(const cse_model* m) { return (m && m->lowered) ? m->plan.logits_off : -1; }
int64_t cse_model_probs_offset(const cse_model* m) { return (m && m->lowered) ? m->plan.probs_off : -1; }

int cse_model_finalize(cse_model* m, void* d_shared_workspace, size_t shared_workspace_bytes) {
  CSE_REQUIRE(m, "model_finalize: NULL model");
  if (m->finalized) { set_error("model_finalize: called twice"); return CSE_ERR_STATE; }
  int rc = cse_model_lower(m);
  if (rc) return rc;
  const size_t ws = (size_t)m->plan.workspace_bytes;
  if (d_shared_workspace) {          // members that run back to back may share one activation arena
    CSE_REQUIRE(shared_workspace_bytes >= ws && ((uintptr_t)d_shared_workspace % 1024) == 0,
                "model_finalize: shared workspace of %zu bytes is too small (%zu needed) or not 1024-byte aligned",
                shared_workspace_bytes, ws);
    m->d_workspace = d_shared_workspace;
  } else {
    CSE_CUDA(cudaMalloc(&m->d_workspace, ws));
    m->own_workspace = true;
  }
  CSE_CUDA(cudaMalloc(&m->d_weights, m->plan.arena.size()));
  CSE_CUDA(cudaMemcpy(m->d_weights, m->plan.arena.data(), m->plan.arena.size(), cudaMemcpyHostToDevice));
  if ((rc = cse_plan_create(&m->exec, m->max_batch, m->nb_classes))) return rc;
  for (auto& op : m->plan.ops)
    if ((rc = cse_plan_add_op(m->exec, &op))) return rc;
  if ((rc = cse_plan_finalize(m->exec, m->d_workspace, ws, m->d_weights, m->plan.arena.size(), m->plan.logits_off, m->plan.probs_off)))
    return rc;
  std::vector<uint8_t>().swap(m->plan.arena);
  m->finalized = true;
  return CSE_OK;
}

int cse_model_forward(cse_model* m, const void* d_rgb, const void* d_flow, int n, float* d_logits, float* d_probs, void* stream) {
  CSE_REQUIRE(m, "model_forward: NULL model");
  if (!m->finalized) { set_error("model_forward: model not finalized"); return CSE_ERR_STATE; }
  return cse_plan_run(m->exec, (const uint8_t*)d_rgb, (const uint8_t*)d_flow, n, d_logits, d_probs, stream);
}

int cse_model_forward_shared_input(cse_model* m, const void* d_rgb, const void* d_flow, int n, float* d_logits, float* d_probs,
                                   void* stream) {
  CSE_REQUIRE(m, "model_forward: NULL model");
  if (!m->finalized) { set_error("model_forward: model not finalized"); return CSE_ERR_STATE; }
  return cse_plan_run_from(m->exec, (const uint8_t*)d_rgb, (const uint8_t*)d_flow, n, cse_plan_num_input_ops(m->exec), d_logits,
                           d_probs, stream);
}

void cse_model_destroy(cse_model* m) {
  if (!m) return;
  if (m->exec) cse_plan_destroy(m->exec);
  if (m->d_weights) cudaFree(m->d_weights);
  if (m->own_workspace && m->d_workspace) cudaFree(m->d_workspace);
  delete m;
}

}  // extern "C"
