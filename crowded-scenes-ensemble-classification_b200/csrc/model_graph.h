// Layer graphs of the reference's member architectures, in C++ (host only).
//
// One node per Keras layer of the reference model (weight-less ones included), because the member's
// `*_weights.hdf5` file is read positionally, in Keras' `model.layers` order (model.load_weights(by_name=False),
// train.py:1731-1769), and that order depends on every layer.  The builders restate, layer by layer,
//   C3D        ConvNets3D                                   train.py:1224-1273
//   I3D        conv3d_bn / Inception_architecture           train.py:615-670, 1013-1219, 837-841
//   TwoStream  TwoStream_Inception_Inflated3d               train.py:857-1011 (999-1009)
//   R3D-N      Resnet3DBuilder.build / basic_block / bottleneck / _shortcut3d   train.py:1278-1559
// and are the native twin of cse_b200/graph.py (tests/test_model_api.py holds the two to the same plans).
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace cse {
namespace mdl {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SOFTMAX = 2, ACT_SIGMOID = 3 };

struct WeightSpec {
  std::string name;            // "<layer>/kernel:0" ...
  std::vector<int> shape;
};

struct Node {
  std::string name, op;        // input|conv3d|bn|relu|dropout|add|concat|maxpool|avgpool|zeropad|flatten|dense
  std::vector<std::string> inputs;
  int filters = 0, units = 0;
  int k[3] = {1, 1, 1}, s[3] = {1, 1, 1};
  bool same = false, use_bias = false, bn_scale = true;
  int act = ACT_NONE;
  int pb[3] = {0, 0, 0}, pa[3] = {0, 0, 0};
  int pads[3][2] = {{0, 0}, {0, 0}, {0, 0}};     // zeropad
  std::vector<int> out_shape;                  // (D,H,W,C) or (F)
  std::vector<WeightSpec> weights;             // Keras order
};

inline void same_pads(int size, int k, int s, int* out, int* before, int* after) {
  // TF 'SAME': out = ceil(in/s); total = max((out-1)*s + k - in, 0); before = total / 2 (extra padding at the end)
  *out = (size + s - 1) / s;
  const int total = std::max((*out - 1) * s + k - size, 0);
  *before = total / 2;
  *after = total - *before;
}

struct Graph {
  std::string name;
  bool sequential = false;
  std::vector<Node> nodes;                     // insertion order
  std::map<std::string, int> index;
  std::vector<std::string> inputs;
  std::string output;
  std::map<std::string, int> uid;

  std::string autoname(const std::string& prefix) {      // Keras auto-naming: <prefix>_<n>, counters per prefix
    int n = ++uid[prefix];
    return prefix + "_" + std::to_string(n);
  }
  const Node& at(const std::string& n) const {
    auto it = index.find(n);
    if (it == index.end()) throw std::runtime_error("unknown layer " + n);
    return nodes[it->second];
  }
  const std::vector<int>& shape(const std::string& n) const { return at(n).out_shape; }
  std::string add(Node node) {
    if (index.count(node.name)) throw std::runtime_error("duplicate layer name " + node.name);
    index[node.name] = (int)nodes.size();
    nodes.push_back(std::move(node));
    return nodes.back().name;
  }
  void window(const std::vector<int>& in, Node& n) const {
    n.out_shape.assign(3, 0);
    for (int i = 0; i < 3; ++i) {
      int o, b = 0, a = 0;
      if (n.same) same_pads(in[i], n.k[i], n.s[i], &o, &b, &a);
      else o = (in[i] - n.k[i]) / n.s[i] + 1;
      if (o <= 0) throw std::runtime_error("window does not fit the input of " + n.name);
      n.out_shape[i] = o; n.pb[i] = b; n.pa[i] = a;
    }
  }
  std::string input(const std::vector<int>& shp, const std::string& name) {
    Node n; n.name = name; n.op = "input"; n.out_shape = shp;
    inputs.push_back(name);
    return add(n);
  }
  std::string conv3d(const std::string& x, int filters, const int k[3], const int s[3], bool same, bool use_bias, int act,
                     std::string name = "") {
    if (name.empty()) name = autoname("conv3d");
    Node n; n.name = name; n.op = "conv3d"; n.inputs = {x}; n.filters = filters; n.same = same; n.use_bias = use_bias; n.act = act;
    for (int i = 0; i < 3; ++i) { n.k[i] = k[i]; n.s[i] = s[i]; }
    const std::vector<int>& in = shape(x);
    window(in, n);
    n.out_shape.push_back(filters);
    n.weights.push_back({name + "/kernel:0", {k[0], k[1], k[2], in[3], filters}});
    if (use_bias) n.weights.push_back({name + "/bias:0", {filters}});
    return add(n);
  }
  std::string bn(const std::string& x, bool scale, std::string name = "") {
    if (name.empty()) name = autoname("batch_normalization");
    Node n; n.name = name; n.op = "bn"; n.inputs = {x}; n.bn_scale = scale; n.out_shape = shape(x);
    const int c = n.out_shape.back();
    if (scale) n.weights.push_back({name + "/gamma:0", {c}});
    n.weights.push_back({name + "/beta:0", {c}});
    n.weights.push_back({name + "/moving_mean:0", {c}});
    n.weights.push_back({name + "/moving_variance:0", {c}});
    return add(n);
  }
  std::string simple(const std::string& op, const std::string& prefix, const std::string& x, std::string name = "") {
    if (name.empty()) name = autoname(prefix);
    Node n; n.name = name; n.op = op; n.inputs = {x}; n.out_shape = shape(x);
    return add(n);
  }
  std::string relu(const std::string& x, std::string name = "") { return simple("relu", "activation", x, name); }
  std::string dropout(const std::string& x, std::string name = "") { return simple("dropout", "dropout", x, name); }
  std::string add_(const std::vector<std::string>& xs, std::string name = "") {
    if (name.empty()) name = autoname("add");
    Node n; n.name = name; n.op = "add"; n.inputs = xs; n.out_shape = shape(xs[0]);
    for (auto& x : xs) if (shape(x) != n.out_shape) throw std::runtime_error("add: shape mismatch");
    return add(n);
  }
  std::string concat(const std::vector<std::string>& xs, std::string name = "") {
    if (name.empty()) name = autoname("concatenate");
    Node n; n.name = name; n.op = "concat"; n.inputs = xs; n.out_shape = shape(xs[0]);
    int ctot = 0;
    for (auto& x : xs) ctot += shape(x).back();
    n.out_shape.back() = ctot;
    return add(n);
  }
  std::string pool(const std::string& op, const std::string& prefix, const std::string& x, const int k[3], const int s[3],
                   bool same, std::string name = "") {
    if (name.empty()) name = autoname(prefix);
    Node n; n.name = name; n.op = op; n.inputs = {x}; n.same = same;
    for (int i = 0; i < 3; ++i) { n.k[i] = k[i]; n.s[i] = s[i]; }
    const std::vector<int>& in = shape(x);
    window(in, n);
    n.out_shape.push_back(in[3]);
    return add(n);
  }
  std::string maxpool(const std::string& x, const int k[3], const int s[3], bool same, std::string name = "") {
    return pool("maxpool", "max_pooling3d", x, k, s, same, name);
  }
  std::string avgpool(const std::string& x, const int k[3], const int s[3], std::string name = "") {
    return pool("avgpool", "average_pooling3d", x, k, s, false, name);
  }
  std::string zeropad(const std::string& x, const int pads[3][2], std::string name = "") {
    if (name.empty()) name = autoname("zero_padding3d");
    Node n; n.name = name; n.op = "zeropad"; n.inputs = {x};
    const std::vector<int>& in = shape(x);
    n.out_shape = in;
    for (int i = 0; i < 3; ++i) { n.pads[i][0] = pads[i][0]; n.pads[i][1] = pads[i][1]; n.out_shape[i] += pads[i][0] + pads[i][1]; }
    return add(n);
  }
  std::string flatten(const std::string& x, std::string name = "") {
    if (name.empty()) name = autoname("flatten");
    Node n; n.name = name; n.op = "flatten"; n.inputs = {x};
    long long f = 1;
    for (int v : shape(x)) f *= v;
    n.out_shape = {(int)f};
    return add(n);
  }
  std::string dense(const std::string& x, int units, int act, std::string name = "") {
    if (name.empty()) name = autoname("dense");
    Node n; n.name = name; n.op = "dense"; n.inputs = {x}; n.units = units; n.act = act; n.out_shape = {units};
    n.weights.push_back({name + "/kernel:0", {shape(x)[0], units}});
    n.weights.push_back({name + "/bias:0", {units}});
    return add(n);
  }

  // model.layers order.  Sequential: insertion order.  Functional (keras/engine/network.py, 2.2.4): DFS from the
  // output assigns layer_index in first-visit order (inbound layers in call order); depth = longest distance to the
  // output; layers sorted by depth descending, ties by layer_index.
  std::vector<std::string> keras_layer_order() const {
    std::vector<std::string> order;
    if (sequential) {
      for (auto& n : nodes) order.push_back(n.name);
      return order;
    }
    std::map<std::string, int> layer_index, depth;
    std::vector<std::string> post;
    std::set<std::string> seen;
    std::vector<std::pair<std::string, int>> stack;
    stack.push_back({output, 0});
    while (!stack.empty()) {
      auto [name, i] = stack.back();
      stack.pop_back();
      const Node& node = at(name);
      if (i == 0) {
        if (seen.count(name)) continue;
        if (!layer_index.count(name)) { int li = (int)layer_index.size(); layer_index[name] = li; }
      }
      if (i < (int)node.inputs.size()) {
        stack.push_back({name, i + 1});
        const std::string& nxt = node.inputs[i];
        if (!seen.count(nxt)) stack.push_back({nxt, 0});
      } else if (!seen.count(name)) {
        seen.insert(name);
        post.push_back(name);
      }
    }
    for (auto it = post.rbegin(); it != post.rend(); ++it) {
      if (!depth.count(*it)) depth[*it] = 0;
      const int d = depth[*it];
      for (auto& inp : at(*it).inputs) depth[inp] = std::max(depth.count(inp) ? depth[inp] : 0, d + 1);
    }
    order = post;
    std::stable_sort(order.begin(), order.end(), [&](const std::string& a, const std::string& b) {
      if (depth[a] != depth[b]) return depth[a] > depth[b];
      return layer_index[a] < layer_index[b];
    });
    return order;
  }
  std::vector<std::string> weighted_layers() const {
    std::vector<std::string> out;
    for (auto& n : keras_layer_order()) if (!at(n).weights.empty()) out.push_back(n);
    return out;
  }
};

// ----------------------------------------------------------------------------- builders
static const int K1[3] = {1, 1, 1}, K3[3] = {3, 3, 3}, K7[3] = {7, 7, 7}, S1[3] = {1, 1, 1}, S2[3] = {2, 2, 2};

inline Graph build_c3d(const std::vector<int>& in, int classes) {
  Graph g; g.name = "C3D"; g.sequential = true;
  const int p122[3] = {1, 2, 2}, p222[3] = {2, 2, 2};
  std::string x = g.input(in, "conv1_input");
  x = g.conv3d(x, 64, K3, S1, true, true, ACT_RELU, "conv1");
  x = g.maxpool(x, p122, p122, false, "pool1");
  x = g.conv3d(x, 128, K3, S1, true, true, ACT_RELU, "conv2");
  x = g.maxpool(x, p222, p222, false, "pool2");
  x = g.conv3d(x, 256, K3, S1, true, true, ACT_RELU, "conv3a");
  x = g.conv3d(x, 256, K3, S1, true, true, ACT_RELU, "conv3b");
  x = g.maxpool(x, p222, p222, false, "pool3");
  x = g.conv3d(x, 512, K3, S1, true, true, ACT_RELU, "conv4a");
  x = g.conv3d(x, 512, K3, S1, true, true, ACT_RELU, "conv4b");
  x = g.maxpool(x, p222, p222, false, "pool4");
  x = g.conv3d(x, 512, K3, S1, true, true, ACT_RELU, "conv5a");
  x = g.conv3d(x, 512, K3, S1, true, true, ACT_RELU, "conv5b");
  const int zp[3][2] = {{0, 0}, {0, 1}, {0, 1}};
  x = g.zeropad(x, zp, "zeropad5");
  x = g.maxpool(x, p222, p222, false, "pool5");
  x = g.flatten(x);
  x = g.dense(x, 4096, ACT_RELU, "fc6");
  x = g.dropout(x);
  x = g.dense(x, 4096, ACT_RELU, "fc7");
  x = g.dropout(x);
  g.output = g.dense(x, classes, ACT_SOFTMAX, "fc8");
  return g;
}

inline std::string conv3d_bn(Graph& g, const std::string& x, int filters, const int k[3], const int s[3], const std::string& name) {
  // conv3d_bn (train.py:646-668): Conv3D(no bias,'same') '<name>_conv' -> BN(scale=False) '<name>_bn' -> ReLU '<name>'
  std::string y = g.conv3d(x, filters, k, s, true, false, ACT_NONE, name + "_conv");
  y = g.bn(y, false, name + "_bn");
  return g.relu(y, name);
}

inline std::string i3d_tower(Graph& g, std::string x, const std::string& ext) {
  struct Mixed { const char* tag; int f[6]; };
  static const Mixed blocks[] = {
      {"3b", {64, 96, 128, 16, 32, 32}}, {"3c", {128, 128, 192, 32, 96, 64}}, {"POOL4a", {0}},
      {"4b", {192, 96, 208, 16, 48, 64}}, {"4c", {160, 112, 224, 24, 64, 64}}, {"4d", {128, 128, 256, 24, 64, 64}},
      {"4e", {112, 144, 288, 32, 64, 64}}, {"4f", {256, 160, 320, 32, 128, 128}}, {"POOL5a", {0}},
      {"5b", {256, 160, 320, 32, 128, 128}}, {"5c", {384, 192, 384, 48, 128, 128}}};
  const int p133[3] = {1, 3, 3}, s122[3] = {1, 2, 2}, p222[3] = {2, 2, 2};
  x = conv3d_bn(g, x, 64, K7, S2, "Conv3d_1a_7x7" + ext);
  x = g.maxpool(x, p133, s122, true, "MaxPool2d_2a_3x3" + ext);
  x = conv3d_bn(g, x, 64, K1, S1, "Conv3d_2b_1x1" + ext);
  x = conv3d_bn(g, x, 192, K3, S1, "Conv3d_2c_3x3" + ext);
  x = g.maxpool(x, p133, s122, true, "MaxPool2d_3a_3x3" + ext);
  for (const Mixed& m : blocks) {
    const std::string tag = m.tag;
    if (tag == "POOL4a") { x = g.maxpool(x, K3, S2, true, "MaxPool2d_4a_3x3" + ext); continue; }
    if (tag == "POOL5a") { x = g.maxpool(x, p222, p222, true, "MaxPool2d_5a_2x2" + ext); continue; }
    std::string b0 = conv3d_bn(g, x, m.f[0], K1, S1, "Conv3d_" + tag + "_0a_1x1" + ext);
    std::string b1 = conv3d_bn(g, x, m.f[1], K1, S1, "Conv3d_" + tag + "_1a_1x1" + ext);
    b1 = conv3d_bn(g, b1, m.f[2], K3, S1, "Conv3d_" + tag + "_1b_3x3" + ext);
    std::string b2 = conv3d_bn(g, x, m.f[3], K1, S1, "Conv3d_" + tag + "_2a_1x1" + ext);
    b2 = conv3d_bn(g, b2, m.f[4], K3, S1, "Conv3d_" + tag + "_2b_3x3" + ext);
    std::string b3 = g.maxpool(x, K3, S1, true, "MaxPool2d_" + tag + "_3a_3x3" + ext);
    b3 = conv3d_bn(g, b3, m.f[5], K1, S1, "Conv3d_" + tag + "_3b_1x1" + ext);
    x = g.concat({b0, b1, b2, b3}, "Mixed_" + tag + ext);
  }
  const std::vector<int>& shp = g.shape(x);
  const int kavg[3] = {2, shp[1], shp[2]};
  return g.avgpool(x, kavg, S1, "global_avg_pool" + ext);
}

inline Graph build_i3d(const std::vector<int>& in, int classes) {
  Graph g; g.name = "I3D";
  std::string x = g.input(in, "input_1");
  x = i3d_tower(g, x, "_rgb");           // type='rgb' is hard-coded for the single stream (train.py:766)
  x = g.flatten(x);
  g.output = g.dense(x, classes, ACT_SOFTMAX, "predictions");
  return g;
}

inline Graph build_twostream(const std::vector<int>& in, int classes) {
  // flow tower constructed first (train.py:919), rgb second (:927); features concatenated [rgb, flow] (:1006);
  // model inputs [rgb, flow] (:1009)
  Graph g; g.name = "TWOSTREAM_I3D";
  std::string rgb = g.input({in[0], in[1], in[2], 3}, "input_1");
  std::string flow = g.input({in[0], in[1], in[2], 2}, "input_2");
  std::string y = i3d_tower(g, flow, "_flow");
  std::string x = i3d_tower(g, rgb, "_rgb");
  x = g.flatten(x);
  y = g.flatten(y);
  std::string z = g.concat({x, y});
  g.output = g.dense(z, classes, ACT_SOFTMAX, "predictions");
  g.inputs = {rgb, flow};
  return g;
}

inline std::string bn_relu(Graph& g, const std::string& x) { return g.relu(g.bn(x, true)); }
inline std::string bn_relu_conv(Graph& g, const std::string& x, int filters, const int k[3], const int s[3]) {
  return g.conv3d(bn_relu(g, x), filters, k, s, true, true, ACT_NONE);
}
inline std::string shortcut(Graph& g, const std::string& x, const std::string& residual) {
  // _shortcut3d (train.py:1324-1346): 1x1x1 'valid' strided conv on the RAW block input when dims or channels differ
  const std::vector<int> xi = g.shape(x), xr = g.shape(residual);
  int st[3];
  bool strided = false;
  for (int i = 0; i < 3; ++i) { st[i] = (xi[i] + xr[i] - 1) / xr[i]; strided = strided || st[i] > 1; }
  std::string sc = x;
  if (strided || xi[3] != xr[3]) sc = g.conv3d(x, xr[3], K1, st, false, true, ACT_NONE);
  return g.add_({sc, residual});
}
inline std::string basic_block(Graph& g, const std::string& x, int filters, const int s[3], bool first_of_first) {
  std::string c1 = first_of_first ? g.conv3d(x, filters, K3, s, true, true, ACT_NONE) : bn_relu_conv(g, x, filters, K3, s);
  std::string res = bn_relu_conv(g, c1, filters, K3, S1);
  return shortcut(g, x, res);
}
inline std::string bottleneck(Graph& g, const std::string& x, int filters, const int s[3], bool first_of_first) {
  std::string c1 = first_of_first ? g.conv3d(x, filters, K1, s, true, true, ACT_NONE) : bn_relu_conv(g, x, filters, K1, s);
  std::string c3 = bn_relu_conv(g, c1, filters, K3, S1);
  std::string res = bn_relu_conv(g, c3, filters * 4, K1, S1);
  return shortcut(g, x, res);
}

inline bool r3d_repetitions(const std::string& mt, bool* basic, std::vector<int>* reps) {
  if (mt == "R3D_18") { *basic = true; *reps = {2, 2, 2, 2}; return true; }
  if (mt == "R3D_34") { *basic = true; *reps = {3, 4, 6, 3}; return true; }
  if (mt == "R3D_50") { *basic = false; *reps = {3, 4, 6, 3}; return true; }
  if (mt == "R3D_101") { *basic = false; *reps = {3, 4, 23, 3}; return true; }
  if (mt == "R3D_152") { *basic = false; *reps = {3, 8, 36, 3}; return true; }
  return false;
}

inline Graph build_r3d(const std::string& mt, const std::vector<int>& in, int classes) {
  bool basic = true;
  std::vector<int> reps;
  if (!r3d_repetitions(mt, &basic, &reps)) throw std::runtime_error("Unknown model " + mt);
  Graph g; g.name = mt;
  std::string x = g.input(in, "input_1");
  x = g.conv3d(x, 64, K7, S2, true, true, ACT_NONE);           // _conv_bn_relu3D
  x = bn_relu(g, x);
  x = g.maxpool(x, K3, S2, true);
  int filters = 64;
  for (size_t i = 0; i < reps.size(); ++i) {
    for (int j = 0; j < reps[i]; ++j) {
      const int* st = (j == 0 && i != 0) ? S2 : S1;
      x = basic ? basic_block(g, x, filters, st, i == 0 && j == 0) : bottleneck(g, x, filters, st, i == 0 && j == 0);
    }
    filters *= 2;
  }
  const std::vector<int> bs = g.shape(x);
  x = bn_relu(g, x);
  const int kavg[3] = {bs[0], bs[1], bs[2]};
  x = g.avgpool(x, kavg, S1);
  x = g.flatten(x);
  g.output = g.dense(x, classes, classes > 1 ? ACT_SOFTMAX : ACT_SIGMOID);
  return g;
}

inline Graph build_model_graph(const std::string& mt, const std::vector<int>& in, int classes) {
  if (mt == "C3D") return build_c3d(in, classes);
  if (mt == "I3D") return build_i3d(in, classes);
  if (mt == "TWOSTREAM_I3D") return build_twostream(in, classes);
  return build_r3d(mt, in, classes);
}

}  // namespace mdl
}  // namespace cse
