// Full-network executor: the graph of RootNetwithRegInt.forward (lib/models/full_net.py:262-466) as a flat op list,
// built once from the reference-format state dict, planned per batch size (liveness-based workspace, no allocation on
// the timed path) and replayed as one CUDA graph.
//
//   DepthNet      rootnet_backbone (HRNet-W32, HRnet.py:499-570) -> depth_layer -> gamma*k/1000   full_net.py:289-342
//   keypoints     reg_backbone (ResNet-50, Resnet.py:57-68) -> avgpool, 3x deconv + 1x1 -> integral   348-360
//                 or reg_backbone (HRNet-W32 with heatmap conv) -> integral                          361-364
//   heads         4x linear refinement of pose and rot6d                                          376-444
//   kinematics    FK (+ re-rooting) and both pinhole projections                                  447-450, function.py:140
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.h"

struct hrp_fk;  // fk_project.cu

namespace hrp {

enum OpKind { OP_STEM, OP_CONV, OP_MAXPOOL, OP_FUSE, OP_AVGPOOL, OP_DEPTH, OP_HEADS, OP_JOINTMAP, OP_SOFTARGMAX, OP_FK, OP_STEM_PACK, OP_BLOCK, OP_CHAIN };
enum { CLS_CONV_TC = 0, CLS_CONV_F32 = 1, CLS_STEM = 2, CLS_ELEM = 3, CLS_HEADS = 4, CLS_SOFTARGMAX = 5, CLS_FK = 6 };
enum TKind { T_WS = 0, T_XREG, T_XROOT, T_KVAL, T_KMAT, T_FIELD, T_CONST, T_INITP, T_INITR, T_FLAGS };

struct TensorInfo {
  int64_t elems = 0;     // per frame
  int esize = 4;
  TKind kind = T_WS;
  int field = -1;
  const void* cptr = nullptr;  // T_CONST: device pointer (weights arena), batch-independent
  int H = 0, W = 0, C = 0;
  int last_use = -1, first_def = 1 << 30;
  bool keep = false;
  int home_lane = 0;     // lane of the defining op: the tensor lives in that lane's arena
  bool cross = false;    // touched by an op of another lane: never recycled (no cross-stream write-after-read hazards)
  std::string name;
};

struct Layer {            // device-resident packed conv / linear
  float* w = nullptr;       // fp32 family: [K][Cout]
  void* w_tc = nullptr;     // tensor-core families: pack_conv_tc image
  float* bias = nullptr;
  int Cout = 0, Cin = 0, KH = 0, KW = 0;
};

struct OpDesc {
  OpKind kind;
  int cls;
  int in = -1, res = -1, out = -1, in2 = -1, in3 = -1, in4 = -1;
  int out2 = -1, out3 = -1, out4 = -1, out5 = -1;
  int layer = -1, layer2 = -1;   // layer2: second conv of a fused BasicBlock (OP_BLOCK)
  int chain[8] = {-1, -1, -1, -1, -1, -1, -1, -1}, n_chain = 0;   // OP_CHAIN: the convs of a fused branch (conv_chain.cu)
  int Hi = 1, Wi = 1, Cin = 0, Ho = 1, Wo = 1, Cout = 0, KH = 1, KW = 1, stride = 1, pad_h = 0, pad_w = 0;
  int out_sy = 1, out_sx = 1, out_oy = 0, out_ox = 0, Ho_full = 1, Wo_full = 1, relu = 0, out_nchw = 0;
  int res_after_act = 0;
  int x3 = 0;            // TF32 families: this layer runs as 3xTF32 (hi/lo operand split, conv_tc.cu) on full-fp32 activations
  int stem_tc = 0;       // OP_CONV over the packed stem image (custom TMA view); KH = real kernel size, pad_h = real padding
  int same[4] = {-1, -1, -1, -1}, low[3] = {-1, -1, -1}, shift[3] = {0, 0, 0}, n_same = 0, n_low = 0;
  int ld = 0, coff = 0, state_stride = 0, dof = 0, N = 0;
  const float* wptr = nullptr;   // small head weights
  const float* bptr = nullptr;
  double flops = 0.0;    // per frame
  int lane = 0;          // execution lane (stream of the captured graph); ops of one lane run in list order
};

constexpr int kMaxLanes = 12;

struct Plan {
  int B = 0;
  size_t ws_bytes = 0;
  char* ws = nullptr;
  std::vector<size_t> off;
  // static staging the graph reads / writes (the images are NOT staged: the ops that read them run ahead of the graph
  // on the caller's pointers, see run_ops' `part`)
  size_t io_kv = 0, io_K = 0, io_out = 0, io_initp = 0, io_initr = 0, io_flags = 0, sa_ws = 0, sa_ws_bytes = 0;
  size_t io_xf32 = 0;              // fp32 family only: 2 x [B,3,256,256] fp32 for uint8 inputs (the direct stem reads fp32)
  cudaGraphExec_t exec = nullptr;
  cudaEvent_t done = nullptr;      // recorded after the last forward that used this plan
  bool used = false;
  int flags_state = -1;            // what io_flags currently holds (bit 0: init_pose override, bit 1: init_rot override)
  int64_t launches = 0;            // kernels of one forward through this plan (pre-graph ops + graph nodes)
  uint64_t stamp = 0;              // last use, for the least-recently-used trim of the plan cache
};

}  // namespace hrp

using namespace hrp;

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

struct WeightSpec {
  std::string name;
  std::vector<int64_t> shape;
  bool optional = false;  // num_batches_tracked
};

struct hrp_handle {
  hrp_config cfg{};
  int device = 0;
  hrp_fk* fk = nullptr;
  int dof = 0, nkpt = 0, ref_kp = 0;
  std::vector<WeightSpec> specs;
  std::unordered_map<std::string, int> spec_index;
  std::unordered_map<std::string, HostTensor> host;
  bool finalized = false;
  bool use_graph = true;
  std::vector<void*> dev_allocs;
  std::vector<Layer> layers;
  std::vector<TensorInfo> tensors;
  std::vector<OpDesc> ops;
  std::map<int, std::unique_ptr<Plan>> plans;      // key = batch * 8 + slot
  int slots = 4;                   // graph path: up to this many plans (workspace + graph) per batch size; a forward that
  int next_slot = 0;               //   arrives on a DIFFERENT stream than its predecessor takes the next one (so the two
  cudaStream_t last_stream = nullptr; int last_B = 0, last_slot_used = 0;   //   overlap); one stream only ever allocates one
  int max_batches = 4;             // distinct batch sizes kept in the plan cache (least recently used is dropped)
  uint64_t clock = 0;
  std::map<int, int> last_slot;    // batch -> slot of the most recent forward (hrp_debug_tensor)
  std::unordered_map<std::string, int> debug;
  cudaStream_t capture_stream = nullptr;
  int n_lanes = 1;
  cudaStream_t lane_stream[hrp::kMaxLanes] = {};   // lanes 1.. of the captured graph (lane 0 is the capture stream)
  std::vector<cudaEvent_t> events;                 // cross-lane edges, created while capturing
  std::vector<std::vector<int>> waits;             // per op: producer ops on other lanes it must wait for
  std::vector<char> signals;                       // per op: some other lane waits for it
  bool use_lanes = true;
  bool sa_fused = false;       // tensor-core families: the heatmap conv reduces its own logits (conv_tc.cu), no logits tensor
  unsigned long long* timeline = nullptr;           // HRP_TIMELINE: 2 stamps per op (device)
  int lane_pct[hrp::kMaxLanes] = {};               // share of the CTA slots a conv of this lane may occupy (graph mode)
  bool lane_pct_auto = true;                       // no explicit setting: small batches (latency-bound use) get twice the share
  int64_t last_launches = 0;
  int t_xreg = -1, t_xroot = -1, t_kval = -1, t_kmat = -1, t_initp = -1, t_initr = -1, t_flags = -1, t_field[HRP_NUM_FIELDS];
  int field_width[HRP_NUM_FIELDS];
};

namespace {

struct DeviceGuard {      // the entry points run on the handle's device and leave the caller's current device as it was
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev); else if (err == cudaSuccess) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define HRP_ON_DEVICE(h)                                                                                   \
  DeviceGuard _guard((h)->device);                                                                         \
  if (_guard.err != cudaSuccess) return fail(HRP_ERR_CUDA, "cudaSetDevice(%d) failed: %s", (h)->device, cudaGetErrorString(_guard.err))

// development aid (HRP_TIMELINE=<csv path>): %globaltimer stamps around every op of the captured graph
__global__ void stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}

// ---------------------------------------------------------------------------------------------------------------------
// expected tensors, in the reference's naming (mirrors holistic-robot-pose-estimation-study_b200/arch.py)
// ---------------------------------------------------------------------------------------------------------------------
struct SpecBuilder {
  std::vector<WeightSpec>& v;
  void add(const std::string& n, std::vector<int64_t> s, bool opt = false) { v.push_back({n, std::move(s), opt}); }
  void bn(const std::string& p, int c) {
    add(p + ".weight", {c}); add(p + ".bias", {c}); add(p + ".running_mean", {c}); add(p + ".running_var", {c});
    add(p + ".num_batches_tracked", {}, true);
  }
  void conv(const std::string& p, int cin, int cout, int k, bool bias = false) {
    add(p + ".weight", {cout, cin, k, k});
    if (bias) add(p + ".bias", {cout});
  }
  void conv_bn(const std::string& pc, const std::string& pb, int cin, int cout, int k, bool bias = false) {
    conv(pc, cin, cout, k, bias); bn(pb, cout);
  }
  void bottleneck(const std::string& p, int cin, int planes, bool down) {
    conv_bn(p + ".conv1", p + ".bn1", cin, planes, 1);
    conv_bn(p + ".conv2", p + ".bn2", planes, planes, 3);
    conv_bn(p + ".conv3", p + ".bn3", planes, planes * 4, 1);
    if (down) conv_bn(p + ".downsample.0", p + ".downsample.1", cin, planes * 4, 1);
  }
  void linear(const std::string& p, int cin, int cout) { add(p + ".weight", {cout, cin}); add(p + ".bias", {cout}); }
};

const int kResnetBlocks[4] = {3, 4, 6, 3};
const int kResnetPlanes[4] = {64, 128, 256, 512};
const int kHrModules[3] = {1, 4, 3};
const int kHrChannels[4] = {32, 64, 128, 256};
const int kHrHead[4] = {32, 64, 128, 256};

std::string S(const char* fmt, ...) {
  char buf[256];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  return buf;
}

void spec_resnet50(SpecBuilder& sb, const std::string& p) {
  sb.conv_bn(p + "conv1", p + "bn1", 3, 64, 7);
  int cin = 64;
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < kResnetBlocks[l]; ++b) {
      sb.bottleneck(S("%slayer%d.%d", p.c_str(), l + 1, b), cin, kResnetPlanes[l], b == 0);
      cin = kResnetPlanes[l] * 4;
    }
}

void spec_hrnet(SpecBuilder& sb, const std::string& p, int hm_channels) {
  sb.conv_bn(p + "conv1", p + "bn1", 3, 64, 3);
  sb.conv_bn(p + "conv2", p + "bn2", 64, 64, 3);
  for (int b = 0; b < 4; ++b) sb.bottleneck(S("%slayer1.%d", p.c_str(), b), b == 0 ? 64 : 256, 64, b == 0);
  for (int st = 0; st < 3; ++st) {
    const int nb = st + 2;
    const std::string t = S("%stransition%d", p.c_str(), st + 1);
    if (st == 0) {
      sb.conv_bn(t + ".0.0", t + ".0.1", 256, 32, 3);
      sb.conv_bn(t + ".1.0.0", t + ".1.0.1", 256, 64, 3);
    } else {
      sb.conv_bn(S("%s.%d.0.0", t.c_str(), nb - 1), S("%s.%d.0.1", t.c_str(), nb - 1), kHrChannels[nb - 2], kHrChannels[nb - 1], 3);
    }
    for (int m = 0; m < kHrModules[st]; ++m) {
      const std::string s = S("%sstage%d.%d", p.c_str(), st + 2, m);
      for (int bi = 0; bi < nb; ++bi)
        for (int k = 0; k < 4; ++k) {
          const std::string q = S("%s.branches.%d.%d", s.c_str(), bi, k);
          sb.conv_bn(q + ".conv1", q + ".bn1", kHrChannels[bi], kHrChannels[bi], 3);
          sb.conv_bn(q + ".conv2", q + ".bn2", kHrChannels[bi], kHrChannels[bi], 3);
        }
      for (int i = 0; i < nb; ++i)
        for (int j = 0; j < nb; ++j) {
          const std::string f = S("%s.fuse_layers.%d.%d", s.c_str(), i, j);
          if (j > i) sb.conv_bn(f + ".0", f + ".1", kHrChannels[j], kHrChannels[i], 1);
          else if (j < i)
            for (int k = 0; k < i - j; ++k)
              sb.conv_bn(S("%s.%d.0", f.c_str(), k), S("%s.%d.1", f.c_str(), k), kHrChannels[j],
                         k == i - j - 1 ? kHrChannels[i] : kHrChannels[j], 3);
        }
    }
  }
  for (int i = 0; i < 4; ++i) sb.bottleneck(S("%sincre_modules.%d.0", p.c_str(), i), kHrChannels[i], kHrHead[i], true);
  for (int i = 0; i < 3; ++i)
    sb.conv_bn(S("%sdownsamp_modules.%d.0", p.c_str(), i), S("%sdownsamp_modules.%d.1", p.c_str(), i), kHrHead[i] * 4, kHrHead[i + 1] * 4, 3, true);
  sb.conv_bn(p + "final_feat_layer.0", p + "final_feat_layer.1", 1024, 2048, 1, true);
  if (hm_channels) sb.conv(p + "final_layer", 32, hm_channels, 1, true);
}

void build_specs(hrp_handle* h) {
  SpecBuilder sb{h->specs};
  const int hm = h->nkpt * 64;
  if (h->cfg.backbone == HRP_BACKBONE_RESNET50) {
    spec_resnet50(sb, "reg_backbone.");
    int cin = 2048;
    for (int i = 0; i < 3; ++i) {
      sb.add(S("deconv_layers.%d.weight", 3 * i), {cin, 256, 4, 4});
      sb.bn(S("deconv_layers.%d", 3 * i + 1), 256);
      cin = 256;
    }
    sb.conv("final_layer", 256, hm, 1, true);
  } else {
    spec_hrnet(sb, "reg_backbone.", hm);
  }
  if (h->cfg.reg_joint_map) {                        // full_net.py:92-101, 240-258
    const int* jd = h->cfg.joint_conv_dim;
    sb.conv_bn("joint_conv_layers.0", "joint_conv_layers.1", 2048, jd[0], 3, true);
    sb.conv_bn("joint_conv_layers.3", "joint_conv_layers.4", jd[0], jd[1], 3, true);
    sb.conv_bn("joint_conv_layers.6", "joint_conv_layers.7", jd[1], jd[2], 3, true);
    sb.conv("joint_final_layer", jd[2], h->dof, 1, true);
  } else {
    sb.linear("fc_pose_1", 2048 + h->dof, 1024);
    sb.linear("fc_pose_2", 1024, 1024);
    sb.linear("decpose", 1024, h->dof);
  }
  if (h->cfg.direct_reg_rot) {                       // full_net.py:110-117
    sb.linear("fc_rot_1", 2048, 1024);
    for (int i = 2; i <= 6; ++i) sb.linear(S("fc_rot_%d", i), 1024, 1024);
  } else {
    sb.linear("fc_rot_1", 2048 + 6, 1024);
    sb.linear("fc_rot_2", 1024, 1024);
  }
  sb.linear("decrot", 1024, 6);
  spec_hrnet(sb, "rootnet_backbone.", 0);
  if (h->cfg.add_fc) {                               // full_net.py:156-163
    sb.linear("depth_fc_d1", 2048, 1024);
    sb.linear("depth_fc_d2", 1024, 512);
    sb.bn("depth_bn", 512);
    sb.linear("depth_fc_u2", 512, 1024);
    sb.linear("depth_fc_u1", 1024, 2048);
  }
  sb.conv("depth_layer", 2048, std::max(1, h->cfg.depth_num), 1, true);
  sb.add("init_pose", {1, h->dof});
  sb.add("init_rot", {1, 6});
  for (size_t i = 0; i < h->specs.size(); ++i) h->spec_index[h->specs[i].name] = (int)i;
}

// ---------------------------------------------------------------------------------------------------------------------
// graph construction
// ---------------------------------------------------------------------------------------------------------------------
struct Tn { int id = -1, H = 0, W = 0, C = 0; };

struct GraphBuilder {
  hrp_handle* h;
  int status = HRP_OK;
  int act_esize = 4;
  int prec = HRP_PREC_FP32;
  int cur_lane = 0;
  int phase_lane0 = 0;   // first of three extra lanes for the deconv phases (0: keep them on the caller's lane)
  bool x3 = false;       // the sub-network being built runs its convs as 3xTF32 (set per backbone in build())
  void push(OpDesc op) { op.lane = cur_lane; op.x3 = x3 ? 1 : 0; h->ops.push_back(op); }
  bool tc() const { return prec != HRP_PREC_FP32; }
  bool two_byte() const { return prec == HRP_PREC_BF16 || prec == HRP_PREC_F16; }   // bf16 / IEEE half activations
  bool tf32() const { return prec == HRP_PREC_TF32 || prec == HRP_PREC_TF32X3; }

  const float* W(const std::string& name) {
    auto it = h->host.find(name);
    if (it == h->host.end()) { if (status == HRP_OK) status = fail(HRP_ERR_WEIGHT, "missing tensor %s", name.c_str()); return nullptr; }
    return it->second.data.data();
  }
  bool has(const std::string& name) { return h->host.count(name) != 0; }

  float* upload(const std::vector<float>& v) {
    float* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(float)) != cudaSuccess) {
      if (status == HRP_OK) status = fail(HRP_ERR_NOMEM, "cudaMalloc of %zu bytes for weights failed", v.size() * sizeof(float));
      cudaGetLastError();
      return nullptr;
    }
    h->dev_allocs.push_back(d);
    if (cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      if (status == HRP_OK) status = fail(HRP_ERR_CUDA, "weight upload failed");
      cudaGetLastError();
    }
    return d;
  }

  void* upload_bytes(const std::vector<uint8_t>& v) {
    void* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(v.size(), 16)) != cudaSuccess) {
      if (status == HRP_OK) status = fail(HRP_ERR_NOMEM, "cudaMalloc of %zu bytes for weights failed", v.size());
      cudaGetLastError();
      return nullptr;
    }
    h->dev_allocs.push_back(d);
    if (cudaMemcpy(d, v.data(), v.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      if (status == HRP_OK) status = fail(HRP_ERR_CUDA, "weight upload failed");
      cudaGetLastError();
    }
    return d;
  }

  // finish a backbone conv layer from its fp32 [K][Cout] matrix: the tensor-core families keep only the swizzled image
  // (`g`: the shape fields of the op that will use the layer -- they decide the operand-row width it is packed for)
  int finish_layer(std::vector<float>& wp, float* dbias, int Cin, int Cout, int KH, int KW, const ConvArgs& g) {
    Layer L; L.Cout = Cout; L.Cin = Cin; L.KH = KH; L.KW = KW; L.bias = dbias;
    if (tc()) {
      const int rb = conv_tc_row_bytes(g, tf32(), nullptr);
      const int mode = tf32() ? (x3 ? 2 : 1) : (prec == HRP_PREC_F16 ? 3 : 0);
      std::vector<uint8_t> img(pack_conv_tc_bytes(KH * KW * Cin, Cout, mode, rb));
      pack_conv_tc(wp.data(), KH * KW * Cin, Cout, mode, rb, img.data());
      L.w_tc = upload_bytes(img);
    } else {
      L.w = upload(wp);
    }
    h->layers.push_back(L);
    return (int)h->layers.size() - 1;
  }

  static ConvArgs shape_of(const OpDesc& o) {
    ConvArgs g{};
    g.B = 1; g.Hi = o.Hi; g.Wi = o.Wi; g.Cin = o.Cin; g.Ho = o.Ho; g.Wo = o.Wo; g.Cout = o.Cout; g.KH = o.KH; g.KW = o.KW;
    g.stride = o.stride; g.pad_h = o.pad_h; g.pad_w = o.pad_w; g.ld_out = o.Cout;
    return g;
  }

  Tn new_tensor(int H, int W_, int C, int esize, const char* name = "") {
    TensorInfo t;
    t.elems = (int64_t)H * W_ * C; t.esize = esize; t.H = H; t.W = W_; t.C = C; t.name = name;
    h->tensors.push_back(t);
    return Tn{(int)h->tensors.size() - 1, H, W_, C};
  }
  int special(TKind k, int64_t elems, int field = -1) {
    TensorInfo t; t.kind = k; t.elems = elems; t.field = field; t.keep = true;
    h->tensors.push_back(t);
    return (int)h->tensors.size() - 1;
  }
  int constant(const float* dptr) {
    TensorInfo t; t.kind = T_CONST; t.cptr = dptr; t.keep = true;
    h->tensors.push_back(t);
    return (int)h->tensors.size() - 1;
  }

  // conv (+bias) (+BN) packed for the fp32 family. wname = "<prefix>.weight"; bn prefix may be empty.
  int make_layer(const std::string& pc, const std::string& pb, int Cin, int Cout, int KH, int KW, const ConvArgs& geom) {
    const float* w = W(pc + ".weight");
    const float* cb = has(pc + ".bias") ? W(pc + ".bias") : nullptr;
    const float *g = nullptr, *b = nullptr, *m = nullptr, *v = nullptr;
    if (!pb.empty()) { g = W(pb + ".weight"); b = W(pb + ".bias"); m = W(pb + ".running_mean"); v = W(pb + ".running_var"); }
    if (status != HRP_OK) return -1;
    std::vector<float> wp((size_t)KH * KW * Cin * Cout), bp(Cout);
    pack_conv_f32(w, cb, g, b, m, v, Cout, Cin, KH, KW, wp.data(), bp.data());
    return finish_layer(wp, upload(bp), Cin, Cout, KH, KW, geom);
  }

  Tn conv(const Tn& x, const std::string& pc, const std::string& pb, int Cout, int k, int stride, int pad, int relu,
          int res = -1, int res_after_act = 0, int out_nchw = 0, int out_esize = -1, const char* name = "") {
    OpDesc op{};
    op.kind = OP_CONV; op.cls = tc() ? CLS_CONV_TC : CLS_CONV_F32;
    op.Hi = x.H; op.Wi = x.W; op.Cin = x.C; op.Cout = Cout; op.KH = op.KW = k; op.stride = stride; op.pad_h = op.pad_w = pad;
    op.Ho = (x.H + 2 * pad - k) / stride + 1; op.Wo = (x.W + 2 * pad - k) / stride + 1;
    op.layer = make_layer(pc, pb, x.C, Cout, k, k, shape_of(op));
    op.Ho_full = op.Ho; op.Wo_full = op.Wo; op.relu = relu; op.res = res; op.res_after_act = res_after_act; op.out_nchw = out_nchw;
    Tn y = new_tensor(op.Ho, op.Wo, Cout, out_esize > 0 ? out_esize : act_esize, name);
    op.in = x.id; op.out = y.id; op.ld = Cout;
    op.flops = 2.0 * op.Ho * op.Wo * Cout * k * k * x.C;
    push(op);
    return y;
  }

  // Tensor-core families: pack the image (zero-padded NHWC4 in the family's operand type), then an implicit GEMM with
  // K = k rows x (8-pixel x 4-channel window) through a custom TMA view (kernels.h, stem_pack_launch).
  Tn stem_tc(int x_ext, const std::string& pc, const std::string& pb, int k, int pad) {
    const float* w = W(pc + ".weight");
    const float *g = W(pb + ".weight"), *b = W(pb + ".bias"), *m = W(pb + ".running_mean"), *v = W(pb + ".running_var");
    Tn packed = new_tensor(STEM_HP, STEM_WP, 4, act_esize);
    Tn y = new_tensor(128, 128, 64, act_esize);
    if (status != HRP_OK) return y;
    std::vector<float> wp((size_t)k * 32 * 64, 0.f), bp(64);
    for (int o = 0; o < 64; ++o) {
      const double sc = (double)g[o] / std::sqrt((double)v[o] + 1e-5);
      bp[o] = (float)((double)b[o] - (double)m[o] * sc);
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < k; ++r)
          for (int q = 0; q < k; ++q)
            wp[((size_t)r * 32 + q * 4 + c) * 64 + o] = (float)((double)w[(((size_t)o * 3 + c) * k + r) * k + q] * sc);
    }
    OpDesc pk{};
    pk.kind = OP_STEM_PACK; pk.cls = CLS_STEM; pk.in = x_ext; pk.out = packed.id;
    push(pk);
    OpDesc op{};
    op.kind = OP_CONV; op.cls = CLS_CONV_TC; op.stem_tc = 1;
    // the GEMM the kernel sees: "Cin" = one 32-element window, KH rows, KW = 1, stride 2 over padded rows
    op.Hi = STEM_HP; op.Wi = 128; op.Cin = 32; op.Cout = 64; op.KH = k; op.KW = 1; op.stride = 2; op.pad_h = pad - STEM_PAD; op.pad_w = 0;
    op.Ho = op.Wo = op.Ho_full = op.Wo_full = 128; op.relu = 1; op.ld = 64;
    op.layer = finish_layer(wp, upload(bp), 32, 64, k, 1, shape_of(op));
    op.out_sy = op.out_sx = 1;
    op.pad_w = pad;                                  // real padding, consumed when the TMA view is built
    op.in = packed.id; op.out = y.id;
    op.flops = 2.0 * 128 * 128 * 64 * k * k * 3;     // the convolution's own FLOPs, not the zero-padded GEMM's
    push(op);
    return y;
  }

  Tn stem(int x_ext, const std::string& pc, const std::string& pb, int k, int pad) {
    if (tc()) return stem_tc(x_ext, pc, pb, k, pad);
    // weights packed [(c*KH + r)*KW + s][64]
    const float* w = W(pc + ".weight");
    const float *g = W(pb + ".weight"), *b = W(pb + ".bias"), *m = W(pb + ".running_mean"), *v = W(pb + ".running_var");
    Tn y = new_tensor(128, 128, 64, act_esize);
    if (status != HRP_OK) return y;
    std::vector<float> wp((size_t)3 * k * k * 64), bp(64);
    for (int o = 0; o < 64; ++o) {
      const double sc = (double)g[o] / std::sqrt((double)v[o] + 1e-5);
      bp[o] = (float)((double)b[o] - (double)m[o] * sc);
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < k; ++r)
          for (int s = 0; s < k; ++s)
            wp[((size_t)(c * k + r) * k + s) * 64 + o] = (float)((double)w[(((size_t)o * 3 + c) * k + r) * k + s] * sc);
    }
    Layer L; L.Cout = 64; L.Cin = 3; L.KH = L.KW = k; L.w = upload(wp); L.bias = upload(bp);
    h->layers.push_back(L);
    OpDesc op{};
    op.kind = OP_STEM; op.cls = CLS_STEM; op.layer = (int)h->layers.size() - 1;
    op.in = x_ext; op.out = y.id; op.Hi = op.Wi = 256; op.Ho = op.Wo = 128; op.KH = op.KW = k; op.pad_h = pad;
    op.flops = 2.0 * 128 * 128 * 64 * k * k * 3;
    push(op);
    return y;
  }

  Tn bottleneck(const Tn& x, const std::string& p, int planes, int stride) {
    Tn y = conv(x, p + ".conv1", p + ".bn1", planes, 1, 1, 0, 1);
    y = conv(y, p + ".conv2", p + ".bn2", planes, 3, stride, 1, 1);
    Tn r = x;
    if (has(p + ".downsample.0.weight")) r = conv(x, p + ".downsample.0", p + ".downsample.1", planes * 4, 1, stride, 0, 0);
    return conv(y, p + ".conv3", p + ".bn3", planes * 4, 1, 1, 0, 1, r.id);   // relu(bn3(conv3) + residual)
  }

  Tn basic(const Tn& x, const std::string& p) {
    BlockArgs probe{};
    probe.B = 1; probe.H = x.H; probe.W = x.W; probe.C = x.C;
    if (two_byte() && conv_block_supported(probe)) {
      // both convs in one launch, the intermediate stays in shared memory (conv_block.cu)
      OpDesc op{};
      op.kind = OP_BLOCK; op.cls = CLS_CONV_TC;
      op.Hi = op.Ho = op.Ho_full = x.H; op.Wi = op.Wo = op.Wo_full = x.W; op.Cin = op.Cout = x.C; op.KH = op.KW = 3; op.stride = 1; op.pad_h = op.pad_w = 1;
      op.relu = 1; op.ld = x.C;
      op.layer = make_layer(p + ".conv1", p + ".bn1", x.C, x.C, 3, 3, shape_of(op));
      op.layer2 = make_layer(p + ".conv2", p + ".bn2", x.C, x.C, 3, 3, shape_of(op));
      Tn y = new_tensor(x.H, x.W, x.C, act_esize);
      op.in = x.id; op.out = y.id;
      op.flops = 2.0 * 2.0 * x.H * x.W * x.C * 9 * x.C;
      push(op);
      return y;
    }
    Tn y = conv(x, p + ".conv1", p + ".bn1", x.C, 3, 1, 1, 1);
    return conv(y, p + ".conv2", p + ".bn2", x.C, 3, 1, 1, 1, x.id);
  }

  // Branch i of an HRNet (its blocks and the fuse ops that produce output i) runs on lane lane0 + i: the branches of a
  // module are independent until the fuse layers exchange them (HRnet.py:247-265), and the low-resolution ones are far
  // too small to fill 148 SMs on their own.
  // The four BasicBlocks of a low-resolution branch as one launch (conv_chain.cu), when the kernel takes the shape.
  bool branch_chain(Tn* x, const std::string& p) {
    ChainArgs probe{};
    probe.B = 1; probe.H = x->H; probe.W = x->W; probe.C = x->C; probe.nconv = 8;
    if (!two_byte() || !(conv_chain_supported(probe) || conv_roll_supported(probe))) return false;
    OpDesc op{};
    op.kind = OP_CHAIN; op.cls = CLS_CONV_TC;
    op.Hi = op.Ho = op.Ho_full = x->H; op.Wi = op.Wo = op.Wo_full = x->W; op.Cin = op.Cout = x->C; op.KH = op.KW = 3; op.stride = 1; op.pad_h = op.pad_w = 1;
    op.relu = 1; op.ld = x->C;
    for (int k = 0; k < 4; ++k) {
      const std::string b = S("%s.%d", p.c_str(), k);
      op.chain[op.n_chain++] = make_layer(b + ".conv1", b + ".bn1", x->C, x->C, 3, 3, shape_of(op));
      op.chain[op.n_chain++] = make_layer(b + ".conv2", b + ".bn2", x->C, x->C, 3, 3, shape_of(op));
    }
    op.layer = op.chain[0];
    Tn y = new_tensor(x->H, x->W, x->C, act_esize);
    op.in = x->id; op.out = y.id;
    // scratch for the layer-by-layer form the executor uses for small batches (one CTA per image: a handful of images
    // is a handful of SMs walking the eight convs serially, slower than eight wide launches)
    op.out2 = new_tensor(x->H, x->W, x->C, act_esize).id;
    op.out3 = new_tensor(x->H, x->W, x->C, act_esize).id;
    op.out4 = new_tensor(x->H, x->W, x->C, act_esize).id;
    op.flops = 8.0 * 2.0 * x->H * x->W * x->C * 9 * x->C;
    push(op);
    *x = y;
    return true;
  }

  std::vector<Tn> hr_module(std::vector<Tn> xs, const std::string& p, int lane0) {
    const int n = (int)xs.size();
    for (int i = 0; i < n; ++i) {
      cur_lane = lane0 + i;
      if (branch_chain(&xs[i], S("%s.branches.%d", p.c_str(), i))) continue;
      for (int k = 0; k < 4; ++k) xs[i] = basic(xs[i], S("%s.branches.%d.%d", p.c_str(), i, k));
    }
    std::vector<Tn> out(n);
    for (int i = 0; i < n; ++i) {
      cur_lane = lane0 + i;
      Tn acc = xs[i];
      const bool has_low = i < n - 1;
      for (int j = 0; j < i; ++j) {  // strided 3x3 chains from higher resolutions; the last conv adds the running sum
        Tn t = xs[j];
        for (int k = 0; k < i - j; ++k) {
          const bool last = k == i - j - 1;
          const std::string f = S("%s.fuse_layers.%d.%d.%d", p.c_str(), i, j, k);
          const int relu = last ? ((j == i - 1 && !has_low) ? 1 : 0) : 1;
          t = conv(t, f + ".0", f + ".1", last ? xs[i].C : xs[j].C, 3, 2, 1, relu, last ? acc.id : -1);
        }
        acc = t;
      }
      if (!has_low) { out[i] = acc; continue; }
      OpDesc op{};
      op.kind = OP_FUSE; op.cls = CLS_ELEM;
      op.same[0] = acc.id; op.n_same = 1;
      for (int j = i + 1; j < n; ++j) {  // 1x1 conv + BN at low resolution, nearest upsample folded into the sum
        const std::string f = S("%s.fuse_layers.%d.%d", p.c_str(), i, j);
        Tn t = conv(xs[j], f + ".0", f + ".1", xs[i].C, 1, 1, 0, 0);
        op.low[op.n_low] = t.id; op.shift[op.n_low] = j - i; ++op.n_low;
      }
      Tn y = new_tensor(xs[i].H, xs[i].W, xs[i].C, act_esize);
      op.out = y.id; op.Ho = y.H; op.Wo = y.W; op.Cout = y.C; op.relu = 1;
      push(op);
      out[i] = y;
    }
    cur_lane = lane0;
    return out;
  }

  // returns feat tensor id (fp32 [2048]); *hm receives the NCHW heatmap tensor when hm_channels > 0
  Tn hrnet(int x_ext, const std::string& p, int hm_channels, Tn* hm, int lane0) {
    cur_lane = lane0;
    Tn x = stem(x_ext, p + "conv1", p + "bn1", 3, 1);
    x = conv(x, p + "conv2", p + "bn2", 64, 3, 2, 1, 1);
    for (int b = 0; b < 4; ++b) x = bottleneck(x, S("%slayer1.%d", p.c_str(), b), 64, 1);
    std::vector<Tn> ys;
    ys.push_back(conv(x, p + "transition1.0.0", p + "transition1.0.1", 32, 3, 1, 1, 1));
    cur_lane = lane0 + 1;
    ys.push_back(conv(x, p + "transition1.1.0.0", p + "transition1.1.0.1", 64, 3, 2, 1, 1));
    ys = hr_module(ys, p + "stage2.0", lane0);
    for (int st = 1; st < 3; ++st) {
      const int nb = st + 2;
      const std::string t = S("%stransition%d.%d.0", p.c_str(), st + 1, nb - 1);
      cur_lane = lane0 + nb - 1;
      ys.push_back(conv(ys.back(), t + ".0", t + ".1", kHrChannels[nb - 1], 3, 2, 1, 1));   // reads y_list[-1], HRnet.py:519
      for (int m = 0; m < kHrModules[st]; ++m) ys = hr_module(ys, S("%sstage%d.%d", p.c_str(), st + 2, m), lane0);
    }
    cur_lane = lane0;
    if (hm_channels) *hm = conv(ys[0], p + "final_layer", "", hm_channels, 1, 1, 0, 0, -1, 0, 1, 4, "logits");
    Tn y = bottleneck(ys[0], p + "incre_modules.0.0", kHrHead[0], 1);
    for (int i = 0; i < 3; ++i) {
      cur_lane = lane0 + i + 1;
      Tn a = bottleneck(ys[i + 1], S("%sincre_modules.%d.0", p.c_str(), i + 1), kHrHead[i + 1], 1);
      const std::string d = S("%sdownsamp_modules.%d", p.c_str(), i);
      y = conv(y, d + ".0", d + ".1", kHrHead[i + 1] * 4, 3, 2, 1, 1, a.id, /*res_after_act=*/1);   // incre(x) + relu(bn(conv(y)))
    }
    y = conv(y, p + "final_feat_layer.0", p + "final_feat_layer.1", 2048, 1, 1, 0, 1);
    return avgpool(y);      // stays on lane0 + 3 (the caller continues there)
  }

  Tn avgpool(const Tn& y) {
    Tn f = new_tensor(1, 1, y.C, 4);
    OpDesc op{};
    op.kind = OP_AVGPOOL; op.cls = CLS_ELEM; op.in = y.id; op.out = f.id; op.Hi = y.H; op.Wi = y.W; op.Cin = y.C;
    push(op);
    return f;
  }

  Tn resnet50(int x_ext, const std::string& p) {
    Tn x = stem(x_ext, p + "conv1", p + "bn1", 7, 3);
    Tn y = new_tensor(64, 64, 64, act_esize);
    OpDesc op{};
    op.kind = OP_MAXPOOL; op.cls = CLS_ELEM; op.in = x.id; op.out = y.id; op.Hi = x.H; op.Wi = x.W; op.Cin = 64;
    push(op);
    x = y;
    for (int l = 0; l < 4; ++l)
      for (int b = 0; b < kResnetBlocks[l]; ++b)
        x = bottleneck(x, S("%slayer%d.%d", p.c_str(), l + 1, b), kResnetPlanes[l], (b == 0 && l > 0) ? 2 : 1);
    return x;
  }

  // ConvTranspose2d(k=4, s=2, p=1) + BN + ReLU as four 2x2 sub-pixel phase convolutions (full_net.py:214-238)
  Tn deconv(const Tn& x, int idx, int Cout) {
    const std::string wn = S("deconv_layers.%d.weight", 3 * idx), pb = S("deconv_layers.%d", 3 * idx + 1);
    const float* w = W(wn);
    const float *g = W(pb + ".weight"), *b = W(pb + ".bias"), *m = W(pb + ".running_mean"), *v = W(pb + ".running_var");
    Tn y = new_tensor(x.H * 2, x.W * 2, Cout, act_esize);
    if (status != HRP_OK) return y;
    const int Cin = x.C;
    std::vector<double> sc(Cout);
    std::vector<float> bp(Cout);
    for (int o = 0; o < Cout; ++o) { sc[o] = (double)g[o] / std::sqrt((double)v[o] + 1e-5); bp[o] = (float)((double)b[o] - (double)m[o] * sc[o]); }
    float* dbias = upload(bp);
    const int lane0 = cur_lane;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        // the four sub-pixel phases are independent GEMMs over the same input: one lane each (phase 0 stays on the
        // caller's lane), joined again by whatever reads y
        if (phase_lane0 > 0 && (py | px)) cur_lane = phase_lane0 + py * 2 + px - 1;
        std::vector<float> wp((size_t)4 * Cin * Cout);
        for (int ty = 0; ty < 2; ++ty)
          for (int tx = 0; tx < 2; ++tx) {
            const int ky = 3 - py - 2 * ty, kx = 3 - px - 2 * tx;
            for (int c = 0; c < Cin; ++c)
              for (int o = 0; o < Cout; ++o)
                wp[((size_t)(ty * 2 + tx) * Cin + c) * Cout + o] = (float)((double)w[(((size_t)c * Cout + o) * 4 + ky) * 4 + kx] * sc[o]);
          }
        OpDesc op{};
        op.kind = OP_CONV; op.cls = tc() ? CLS_CONV_TC : CLS_CONV_F32;
        op.in = x.id; op.out = y.id; op.Hi = x.H; op.Wi = x.W; op.Cin = Cin; op.Cout = Cout; op.KH = op.KW = 2; op.stride = 1;
        op.pad_h = 1 - py; op.pad_w = 1 - px; op.Ho = x.H; op.Wo = x.W;
        op.layer = finish_layer(wp, dbias, Cin, Cout, 2, 2, shape_of(op)); op.out_sy = op.out_sx = 2; op.out_oy = py; op.out_ox = px;
        op.Ho_full = y.H; op.Wo_full = y.W; op.relu = 1; op.ld = Cout;
        op.flops = 2.0 * x.H * x.W * Cout * 4 * Cin;
        push(op);
      }
    cur_lane = lane0;
    return y;
  }

  // The refinement loops are linear (no activation; dropout = identity in eval, full_net.py:376-394, 430-444):
  //   s <- s + Wd (W2 (W1x xf + W1s s + b1) + b2) + bd  =  (I + M) s + A xf + c,
  //   A = Wd W2 W1x, M = Wd W2 W1s, c = Wd (W2 b1 + b2) + bd,
  // so iterate n is  s_n = G_n xf + P_n s_0 + g_n  with  P_n = (I+M)^n, S_n = sum_{j<n} (I+M)^j, G_n = S_n A, g_n = S_n c.
  // Composed here in fp64, evaluated by ONE launch for every iterate of both heads (heads_affine_kernel).
  void heads(const Tn& xf) {
    // reg_joint_map: the joint angles come from joint_map_head (below); the kernel then sees a rotation head only (dof = 0)
    const int dof = h->cfg.reg_joint_map ? 0 : h->dof, F = 2048, Hd = 1024, nit = h->cfg.n_iter, R1 = dof + 6;
    const char* fc1[2] = {"fc_pose_1", "fc_rot_1"};
    const char* fc2[2] = {"fc_pose_2", "fc_rot_2"};
    const char* dec[2] = {"decpose", "decrot"};
    const char* init[2] = {"init_pose", "init_rot"};
    const int sd[2] = {dof, 6}, row0[2] = {0, dof};
    std::vector<float> G((size_t)nit * R1 * F), P((size_t)nit * (dof * dof + 36)), gv((size_t)nit * R1), s0(R1);
    const bool direct = h->cfg.direct_reg_rot != 0, matmul = !direct && h->cfg.rot_iterative_matmul != 0;
    for (int k = 0; k < 2; ++k) {
      if (k == 0 && dof == 0) continue;
      if (k == 1 && direct) { direct_rot(G, P, gv, s0); continue; }
      const int n = sd[k], in1 = F + n;
      const float* w1 = W(std::string(fc1[k]) + ".weight"); const float* b1 = W(std::string(fc1[k]) + ".bias");
      const float* w2 = W(std::string(fc2[k]) + ".weight"); const float* b2 = W(std::string(fc2[k]) + ".bias");
      const float* wd = W(std::string(dec[k]) + ".weight"); const float* bd = W(std::string(dec[k]) + ".bias");
      const float* i0 = W(init[k]);
      if (status != HRP_OK) return;
      for (int j = 0; j < n; ++j) s0[row0[k] + j] = i0[j];
      std::vector<double> T((size_t)n * Hd, 0.0);                       // Wd W2
      for (int j = 0; j < n; ++j)
        for (int m = 0; m < Hd; ++m) {
          const double d = wd[(size_t)j * Hd + m];
          const float* r2 = w2 + (size_t)m * Hd;
          double* t = &T[(size_t)j * Hd];
          for (int q = 0; q < Hd; ++q) t[q] += d * (double)r2[q];
        }
      std::vector<double> A((size_t)n * F, 0.0), M((size_t)n * n, 0.0), c(n, 0.0);
      for (int j = 0; j < n; ++j) {
        double cj = bd[j];
        for (int m = 0; m < Hd; ++m) {
          const double t = T[(size_t)j * Hd + m];
          const float* r1 = w1 + (size_t)m * in1;                       // state columns are LAST (cat([xf, state]))
          double* a = &A[(size_t)j * F];
          for (int q = 0; q < F; ++q) a[q] += t * (double)r1[q];
          for (int q = 0; q < n; ++q) M[(size_t)j * n + q] += t * (double)r1[F + q];
          cj += t * (double)b1[m] + (double)wd[(size_t)j * Hd + m] * (double)b2[m];
        }
        c[j] = cj;
      }
      std::vector<double> E(M), Pn((size_t)n * n, 0.0), Sn((size_t)n * n, 0.0), tmp((size_t)n * n);
      for (int j = 0; j < n; ++j) { E[(size_t)j * n + j] += 1.0; Pn[(size_t)j * n + j] = 1.0; }
      if (k == 1 && matmul) {
        // rot_iterative_matmul: the update is not additive, so nothing composes across iterations. Iterate 0 carries
        // A, c and M (rows of later iterates stay zero); heads_affine_kernel runs the loop itself (rot_matmul mode)
        for (int j = 0; j < n; ++j) {
          const size_t r = (size_t)row0[k] + j;
          for (int q = 0; q < F; ++q) G[r * F + q] = (float)A[(size_t)j * F + q];
          gv[r] = (float)c[j];
          for (int q = 0; q < n; ++q) P[(size_t)dof * dof + (size_t)j * n + q] = (float)M[(size_t)j * n + q];
        }
        continue;
      }
      for (int it = 0; it < nit; ++it) {
        for (size_t q = 0; q < Sn.size(); ++q) Sn[q] += Pn[q];          // S_{it+1} = S_it + P_it
        for (int j = 0; j < n; ++j)                                     // P_{it+1} = E P_it
          for (int q = 0; q < n; ++q) {
            double v = 0.0;
            for (int m = 0; m < n; ++m) v += E[(size_t)j * n + m] * Pn[(size_t)m * n + q];
            tmp[(size_t)j * n + q] = v;
          }
        Pn = tmp;
        for (int j = 0; j < n; ++j) {
          const size_t r = (size_t)it * R1 + row0[k] + j;
          double gj = 0.0;
          std::vector<double> row(F, 0.0);
          for (int m = 0; m < n; ++m) {
            const double sv = Sn[(size_t)j * n + m];
            gj += sv * c[m];
            const double* a = &A[(size_t)m * F];
            for (int q = 0; q < F; ++q) row[q] += sv * a[q];
          }
          for (int q = 0; q < F; ++q) G[r * F + q] = (float)row[q];
          gv[r] = (float)gj;
          float* pr = &P[(size_t)it * (dof * dof + 36) + (k ? dof * dof : 0) + (size_t)j * n];
          for (int q = 0; q < n; ++q) pr[q] = (float)Pn[(size_t)j * n + q];
        }
      }
    }
    Tn iters = new_tensor(1, 1, nit * R1, 4, "head_iters");
    h->tensors[iters.id].keep = true; h->debug["head_iters"] = iters.id;
    OpDesc op{};
    op.kind = OP_HEADS; op.cls = CLS_HEADS; op.in = xf.id; op.in2 = h->t_initp; op.in3 = h->t_initr; op.in4 = h->t_flags;
    op.out = h->t_field[HRP_F_POSE]; op.out2 = h->t_field[HRP_F_ROT]; op.out3 = iters.id;
    op.same[0] = constant(upload(G)); op.same[1] = constant(upload(P)); op.same[2] = constant(upload(gv)); op.same[3] = constant(upload(s0));
    op.n_same = 4; op.Cin = F; op.dof = dof; op.N = nit; op.relu = matmul ? 1 : 0;     // relu: the kernel's rot_matmul mode
    op.flops = 2.0 * nit * R1 * F;
    push(op);
  }

  // direct_reg_rot (full_net.py:395-409): rot = decrot(fc6(fc5(fc4(fc3(fc2(x1))))) + x1), x1 = fc1(xf): seven linear layers and a
  // skip, no activation -> one affine map of xf, the same for every "iterate" and independent of init_rot. Row vectors of
  // decrot are pulled back through the layers in fp64.
  void direct_rot(std::vector<float>& G, std::vector<float>& P, std::vector<float>& gv, std::vector<float>& s0) {
    const int dof = h->cfg.reg_joint_map ? 0 : h->dof, F = 2048, Hd = 1024, nit = h->cfg.n_iter, R1 = dof + 6;
    const float* wd = W("decrot.weight"); const float* bd = W("decrot.bias");
    const float* w1 = W("fc_rot_1.weight"); const float* b1 = W("fc_rot_1.bias");
    const float* i0 = W("init_rot");
    if (status != HRP_OK) return;
    for (int j = 0; j < 6; ++j) s0[dof + j] = i0[j];
    std::vector<double> T((size_t)6 * Hd), acc((size_t)6 * Hd), cst(6);
    for (int j = 0; j < 6; ++j) { cst[j] = bd[j]; for (int m = 0; m < Hd; ++m) T[(size_t)j * Hd + m] = wd[(size_t)j * Hd + m]; }
    for (int l = 6; l >= 2; --l) {                                 // T <- T W_l, constants pick up T b_l on the way
      const float* wl = W(S("fc_rot_%d.weight", l)); const float* bl = W(S("fc_rot_%d.bias", l));
      if (status != HRP_OK) return;
      std::fill(acc.begin(), acc.end(), 0.0);
      for (int j = 0; j < 6; ++j)
        for (int m = 0; m < Hd; ++m) {
          const double t = T[(size_t)j * Hd + m];
          cst[j] += t * (double)bl[m];
          const float* row = wl + (size_t)m * Hd;
          double* a = &acc[(size_t)j * Hd];
          for (int q = 0; q < Hd; ++q) a[q] += t * (double)row[q];
        }
      T = acc;
    }
    for (int j = 0; j < 6; ++j)
      for (int m = 0; m < Hd; ++m) T[(size_t)j * Hd + m] += (double)wd[(size_t)j * Hd + m];      // the skip: x6 + x1
    for (int j = 0; j < 6; ++j) {
      std::vector<double> row(F, 0.0);
      double cj = cst[j];
      for (int m = 0; m < Hd; ++m) {
        const double t = T[(size_t)j * Hd + m];
        cj += t * (double)b1[m];
        const float* r1 = w1 + (size_t)m * F;
        for (int q = 0; q < F; ++q) row[q] += t * (double)r1[q];
      }
      for (int it = 0; it < nit; ++it) {
        const size_t r = (size_t)it * R1 + dof + j;
        for (int q = 0; q < F; ++q) G[r * F + q] = (float)row[q];
        gv[r] = (float)cj;
        for (int q = 0; q < 6; ++q) P[(size_t)it * (dof * dof + 36) + (size_t)dof * dof + (size_t)j * 6 + q] = 0.f;
      }
    }
  }

  // DepthNet head of the add_fc / multi_kp variants (full_net.py:293-330), composed in fp64 around the one nonlinearity:
  //   d1 = W1 f + b1;  z = lrelu(s (W2 d1 + b2 - mu) + beta), s = gamma / sqrt(var + eps);
  //   f3 = (Wu2 z + bu2 + d1) / 2;  f4 = (Wu1 f3 + bu1 + f) / 2;  gamma_d = wd_d . f4 + bd_d
  //   => z = lrelu(Wz f + bz),  gamma_d = Af_d . f + Bz_d . z + c_d   (conv_f32.cu: depth_head_ex_kernel)
  int depth_head_variant(const Tn& img_feat) {
    const int F = 2048, H1 = 1024, Z = h->cfg.add_fc ? 512 : 0, dn = std::max(1, h->cfg.depth_num);
    const float* wd = W("depth_layer.weight"); const float* bd = W("depth_layer.bias");
    if (status != HRP_OK) return status;
    std::vector<float> Af((size_t)dn * F), Bz((size_t)dn * std::max(Z, 1), 0.f), Wz((size_t)std::max(Z, 1) * F, 0.f), bz(std::max(Z, 1), 0.f), cc(dn);
    if (!Z) {
      for (int d = 0; d < dn; ++d) { for (int q = 0; q < F; ++q) Af[(size_t)d * F + q] = wd[(size_t)d * F + q]; cc[d] = bd[d]; }
    } else {
      const float* w1 = W("depth_fc_d1.weight"); const float* b1 = W("depth_fc_d1.bias");
      const float* w2 = W("depth_fc_d2.weight"); const float* b2 = W("depth_fc_d2.bias");
      const float* u2 = W("depth_fc_u2.weight"); const float* c2 = W("depth_fc_u2.bias");
      const float* u1 = W("depth_fc_u1.weight"); const float* c1 = W("depth_fc_u1.bias");
      const float* bw = W("depth_bn.weight"); const float* bb = W("depth_bn.bias");
      const float* bm = W("depth_bn.running_mean"); const float* bv = W("depth_bn.running_var");
      if (status != HRP_OK) return status;
      // Wz = diag(s) W2 W1, bz = s (W2 b1 + b2 - mu) + beta
      for (int r = 0; r < Z; ++r) {
        const double sc = (double)bw[r] / std::sqrt((double)bv[r] + 1e-5);        // nn.BatchNorm1d default eps
        std::vector<double> row(F, 0.0);
        double cst = b2[r];
        for (int m = 0; m < H1; ++m) {
          const double t = w2[(size_t)r * H1 + m];
          cst += t * (double)b1[m];
          const float* r1 = w1 + (size_t)m * F;
          for (int q = 0; q < F; ++q) row[q] += t * (double)r1[q];
        }
        for (int q = 0; q < F; ++q) Wz[(size_t)r * F + q] = (float)(sc * row[q]);
        bz[r] = (float)(sc * (cst - (double)bm[r]) + (double)bb[r]);
      }
      for (int d = 0; d < dn; ++d) {
        const float* wdr = wd + (size_t)d * F;
        std::vector<double> t1(H1, 0.0);                                           // wd_d Wu1  [1024]
        double cst = bd[d];
        for (int q = 0; q < F; ++q) {
          const double t = wdr[q];
          cst += 0.5 * t * (double)c1[q];
          const float* r = u1 + (size_t)q * H1;
          for (int m = 0; m < H1; ++m) t1[m] += t * (double)r[m];
        }
        std::vector<double> af(F, 0.0), bzr(Z, 0.0);
        for (int m = 0; m < H1; ++m) {
          const double t = 0.25 * t1[m];
          cst += t * ((double)c2[m] + (double)b1[m]);
          const float* r2 = u2 + (size_t)m * Z;
          for (int q = 0; q < Z; ++q) bzr[q] += t * (double)r2[q];
          const float* r1 = w1 + (size_t)m * F;
          for (int q = 0; q < F; ++q) af[q] += t * (double)r1[q];
        }
        for (int q = 0; q < F; ++q) Af[(size_t)d * F + q] = (float)(af[q] + 0.5 * (double)wdr[q]);
        for (int q = 0; q < Z; ++q) Bz[(size_t)d * Z + q] = (float)bzr[q];
        cc[d] = (float)cst;
      }
    }
    OpDesc op{};
    op.kind = OP_DEPTH; op.cls = CLS_HEADS; op.in = img_feat.id; op.in2 = h->t_kval; op.out = h->t_field[HRP_F_DEPTH];
    op.out2 = h->cfg.depth_num > 0 ? h->t_field[HRP_F_DEPTHS] : -1;
    op.same[0] = constant(upload(Af)); op.same[1] = constant(upload(Bz)); op.same[2] = constant(upload(Wz)); op.same[3] = constant(upload(bz));
    op.n_same = 4; op.bptr = upload(cc); op.Cin = F; op.Cout = Z; op.N = dn; op.dof = h->cfg.depth_root; op.relu = 1;   // relu: variant kernel
    op.flops = 2.0 * ((double)Z * F + (double)dn * (F + Z));
    push(op);
    return status;
  }

  int build() {
    h->tensors.clear(); h->ops.clear();
    prec = h->cfg.precision;
    act_esize = two_byte() ? 2 : 4;
    {
      const char* e = getenv("HRP_NO_SA_FUSION");
      h->sa_fused = prec != HRP_PREC_FP32 && !(e && atoi(e) != 0);
    }
    const int nk = h->nkpt, dof = h->dof;
    const int fw[HRP_NUM_FIELDS] = {dof, 6, 3, 2, 1, nk * 3, nk * 3, nk * 3, nk * 2, nk * 2, h->cfg.depth_num};
    h->t_xreg = special(T_XREG, 3LL * 256 * 256);
    h->t_xroot = special(T_XROOT, 3LL * 256 * 256);
    h->t_kval = special(T_KVAL, 1);
    h->t_kmat = special(T_KMAT, 9);
    h->t_initp = special(T_INITP, dof);
    h->t_initr = special(T_INITR, 6);
    h->t_flags = special(T_FLAGS, 0);
    for (int f = 0; f < HRP_NUM_FIELDS; ++f) { h->field_width[f] = fw[f]; h->t_field[f] = special(T_FIELD, fw[f], f); }

    // 3xTF32: everywhere in the tf32x3 family; in the tf32 family on the layers of an HRNet-W32 KEYPOINT
    // backbone, whose ~110 sequential single-pass TF32 layers in front of the joint-angle heads measured 1.8e-3 rad against
    // the 1e-3 rad gate (the ResNet-50 keypoint backbone and the DepthNet stay single-pass: 5.4e-4 rad, 0.2 mm)
    const bool x3_all = prec == HRP_PREC_TF32X3;
    bool x3_kp = x3_all || (prec == HRP_PREC_TF32 && h->cfg.backbone == HRP_BACKBONE_HRNET32);
    if (const char* e = getenv("HRP_TF32_X3_KEYPOINT")) x3_kp = x3_all || (prec == HRP_PREC_TF32 && atoi(e) != 0);
    x3 = x3_all;
    // DepthNet: lanes 0-3 (one per HRNet branch); its head ends on lane 3
    Tn img_feat = hrnet(h->t_xroot, "rootnet_backbone.", 0, nullptr, 0);
    h->tensors[img_feat.id].keep = true; h->debug["img_feat"] = img_feat.id;
    if (!h->cfg.add_fc && h->cfg.depth_num == 0) {
      const float* w = W("depth_layer.weight"); const float* b = W("depth_layer.bias");
      if (status != HRP_OK) return status;
      OpDesc op{};
      op.kind = OP_DEPTH; op.cls = CLS_HEADS; op.in = img_feat.id; op.in2 = h->t_kval; op.out = h->t_field[HRP_F_DEPTH];
      op.wptr = upload(std::vector<float>(w, w + 2048)); op.bptr = upload(std::vector<float>(b, b + 1)); op.Cin = 2048;
      op.flops = 2.0 * 2048;
      push(op);
    } else {
      if (depth_head_variant(img_feat) != HRP_OK) return status;
    }
    // keypoint branch: lane 4 (ResNet-50 trunk, deconv head, logits, soft-argmax) or lanes 4-7 (HRNet-W32); the
    // regression heads and FK run on their own lane beside the deconv head
    Tn xf, logits;
    int kp_lane = 4, head_lane = 5;
    x3 = x3_kp;
    if (h->cfg.backbone == HRP_BACKBONE_RESNET50) {
      cur_lane = kp_lane;
      Tn x = resnet50(h->t_xreg, "reg_backbone.");
      cur_lane = head_lane;
      xf = avgpool(x);
      if (h->cfg.reg_joint_map) {                // joint_conv_layers + joint_final_layer + HeatmapIntegralJoint, beside the deconv head
        const int* jd = h->cfg.joint_conv_dim;
        Tn j = conv(x, "joint_conv_layers.0", "joint_conv_layers.1", jd[0], 3, 1, 1, 1);
        j = conv(j, "joint_conv_layers.3", "joint_conv_layers.4", jd[1], 3, 1, 1, 1);
        j = conv(j, "joint_conv_layers.6", "joint_conv_layers.7", jd[2], 3, 1, 1, 1);
        const float* w = W("joint_final_layer.weight"); const float* b = W("joint_final_layer.bias");
        if (status != HRP_OK) return status;
        OpDesc op{};
        op.kind = OP_JOINTMAP; op.cls = CLS_HEADS; op.in = j.id; op.out = h->t_field[HRP_F_POSE]; op.Hi = j.H; op.Wi = j.W; op.Cin = j.C; op.dof = h->dof;
        op.wptr = upload(std::vector<float>(w, w + (size_t)h->dof * j.C)); op.bptr = upload(std::vector<float>(b, b + h->dof));
        op.same[0] = constant(upload(std::vector<float>(h->cfg.joint_bounds, h->cfg.joint_bounds + 2 * h->dof))); op.n_same = 1;
        op.flops = 2.0 * j.H * j.W * j.C * h->dof;
        push(op);
      }
      cur_lane = kp_lane;
      phase_lane0 = head_lane + 1;           // lanes 6, 7, 8
      Tn d = deconv(x, 0, 256);
      d = deconv(d, 1, 256);
      d = deconv(d, 2, 256);
      logits = conv(d, "final_layer", "", nk * 64, 1, 1, 0, 0, -1, 0, /*nchw=*/1, 4, "logits");
    } else {
      xf = hrnet(h->t_xreg, "reg_backbone.", nk * 64, &logits, 4);
      head_lane = 8;                       // xf ends on lane 7, the logits on lane 4
    }
    x3 = false;
    h->n_lanes = head_lane + 1 + (phase_lane0 > 0 ? 3 : 0);
    if (status != HRP_OK) return status;
    cur_lane = kp_lane;
    h->tensors[xf.id].keep = true; h->debug["xf"] = xf.id;
    h->tensors[logits.id].keep = true;
    if (!h->sa_fused) h->debug["logits"] = logits.id;      // fused: the logits never exist outside the conv's accumulators
    {
      OpDesc op{};
      op.kind = OP_SOFTARGMAX; op.cls = CLS_SOFTARGMAX; op.in = logits.id; op.in2 = h->t_kmat; op.in3 = h->t_field[HRP_F_DEPTH];
      op.out = h->t_field[HRP_F_UVD]; op.out2 = h->t_field[HRP_F_XYZ_INT]; op.out3 = h->t_field[HRP_F_ROOT_UV];
      op.out4 = h->t_field[HRP_F_TRANS]; op.out5 = h->t_field[HRP_F_KP2D_INT];
      push(op);
    }
    cur_lane = head_lane;
    heads(xf);
    if (status != HRP_OK) return status;
    {
      OpDesc op{};
      op.kind = OP_FK; op.cls = CLS_FK; op.in = h->t_field[HRP_F_POSE]; op.in2 = h->t_field[HRP_F_ROT]; op.in3 = h->t_field[HRP_F_TRANS];
      op.in4 = h->t_kmat; op.out = h->t_field[HRP_F_XYZ_FK]; op.out2 = h->t_field[HRP_F_KP2D_FK];
      push(op);
    }
    // liveness, lanes and cross-lane dependencies
    // writers of each tensor whose effects a later reader must see: normally the last one; the sub-pixel phases of a
    // transposed conv write disjoint pixels of one tensor, so they do not order among themselves and a reader waits for all
    std::vector<std::vector<int>> last_writers(h->tensors.size());
    h->waits.assign(h->ops.size(), {});
    h->signals.assign(h->ops.size(), 0);
    for (size_t i = 0; i < h->ops.size(); ++i) {
      const OpDesc& o = h->ops[i];
      const int reads[] = {o.in, o.res, o.in2, o.in3, o.in4, o.same[0], o.same[1], o.same[2], o.same[3], o.low[0], o.low[1], o.low[2]};
      const int writes[] = {o.out, o.out2, o.out3, o.out4, o.out5};
      for (int id : writes)
        if (id >= 0 && h->tensors[id].first_def > (int)i) { h->tensors[id].first_def = (int)i; h->tensors[id].home_lane = o.lane; }
      auto touch = [&](int id) {
        TensorInfo& t = h->tensors[id];
        t.last_use = std::max(t.last_use, (int)i);
        t.first_def = std::min(t.first_def, (int)i);
        if (t.kind == T_WS && t.home_lane != o.lane) t.cross = true;
      };
      auto depend = [&](int j) {                 // op i must see op j's effects
        if (j < 0 || h->ops[j].lane == o.lane) return;
        auto& w = h->waits[i];
        if (std::find(w.begin(), w.end(), j) == w.end()) { w.push_back(j); h->signals[j] = 1; }
      };
      for (int id : reads)
        if (id >= 0) { touch(id); for (int j : last_writers[id]) depend(j); }
      for (int id : writes)
        if (id >= 0) {
          touch(id);
          auto& lw = last_writers[id];
          const bool phase = o.kind == OP_CONV && o.out_sy == 2;
          const bool joins = phase && !lw.empty() && h->ops[lw[0]].kind == OP_CONV && h->ops[lw[0]].out_sy == 2 && h->ops[lw[0]].in == o.in;
          if (joins) { lw.push_back((int)i); continue; }
          for (int j : lw) depend(j);                      // write-after-write keeps list order
          lw.assign(1, (int)i);
        }
    }
    return status;
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// planning and execution
// ---------------------------------------------------------------------------------------------------------------------
struct FreeList {   // first-fit with coalescing, offsets in bytes
  std::map<size_t, size_t> free_;  // offset -> size
  size_t top = 0;
  size_t alloc(size_t n) {
    for (auto it = free_.begin(); it != free_.end(); ++it)
      if (it->second >= n) {
        const size_t off = it->first, rest = it->second - n;
        free_.erase(it);
        if (rest) free_[off + n] = rest;
        return off;
      }
    // extend the arena (merging with a trailing free block)
    if (!free_.empty()) {
      auto last = std::prev(free_.end());
      if (last->first + last->second == top) {
        const size_t off = last->first;
        top = off + n;
        free_.erase(last);
        return off;
      }
    }
    const size_t off = top;
    top += n;
    return off;
  }
  void release(size_t off, size_t n) {
    auto it = free_.emplace(off, n).first;
    auto nx = std::next(it);
    if (nx != free_.end() && it->first + it->second == nx->first) { it->second += nx->second; free_.erase(nx); }
    if (it != free_.begin()) {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) { pv->second += it->second; free_.erase(it); }
    }
  }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int64_t record_floats(const hrp_handle* h, int B, int64_t* offs) {
  int64_t o = 0;
  for (int f = 0; f < HRP_NUM_FIELDS; ++f) {
    if (offs) offs[f] = o;
    o += (int64_t)B * h->field_width[f];
    o = (o + 3) & ~3LL;   // 16-byte aligned fields
  }
  if (offs) offs[HRP_NUM_FIELDS] = o;
  return o;
}

void destroy_plan(Plan* p) {
  if (p->used && p->done) cudaEventSynchronize(p->done);
  if (p->exec) cudaGraphExecDestroy(p->exec);
  if (p->done) cudaEventDestroy(p->done);
  if (p->ws) cudaFree(p->ws);
}

// the plan cache keeps at most h->max_batches distinct batch sizes: a caller with ragged batches (detector-driven crops,
// the last batch of an evaluation) must not grow device memory by one workspace per size it ever used
void trim_plans(hrp_handle* h, int keep_B) {
  for (;;) {
    std::map<int, uint64_t> newest;                      // batch -> most recent use over its slots
    for (auto& kv : h->plans) { auto& v = newest[kv.first / 8]; v = std::max(v, kv.second->stamp); }
    if ((int)newest.size() < h->max_batches || (newest.size() == 1 && newest.count(keep_B))) return;
    int victim = -1; uint64_t oldest = ~0ull;
    for (auto& kv : newest) if (kv.first != keep_B && kv.second < oldest) { oldest = kv.second; victim = kv.first; }
    if (victim < 0) return;
    for (auto it = h->plans.begin(); it != h->plans.end();)
      if (it->first / 8 == victim) { destroy_plan(it->second.get()); it = h->plans.erase(it); } else ++it;
    h->last_slot.erase(victim);
  }
}

int make_plan(hrp_handle* h, int B, Plan** out, int slot = 0) {
  auto it = h->plans.find(B * 8 + slot);
  if (it != h->plans.end()) { it->second->stamp = ++h->clock; *out = it->second.get(); return HRP_OK; }
  bool known = false;
  for (auto& kv : h->plans) known |= kv.first / 8 == B;
  if (!known) trim_plans(h, B);
  std::unique_ptr<Plan> p(new Plan());
  p->B = B;
  p->off.assign(h->tensors.size(), 0);
  FreeList fl;
  const size_t A = 256;
  // static I/O staging (graph replays always read/write these)
  p->io_kv = fl.alloc(align_up((size_t)B * 4, A));
  p->io_K = fl.alloc(align_up((size_t)B * 9 * 4, A));
  p->io_initp = fl.alloc(align_up((size_t)B * h->dof * 4, A));
  p->io_initr = fl.alloc(align_up((size_t)B * 6 * 4, A));
  p->io_flags = fl.alloc(A);
  if (h->cfg.precision == HRP_PREC_FP32) p->io_xf32 = fl.alloc(align_up((size_t)2 * B * 3 * 256 * 256 * 4, A));
  p->io_out = fl.alloc(align_up((size_t)record_floats(h, B, nullptr) * 4, A));
  p->sa_ws_bytes = std::max(softargmax_workspace(B, h->nkpt, 64, 64, 64), (size_t)B * h->nkpt * (64 * 64 / 128) * 5 * sizeof(float));
  p->sa_ws = fl.alloc(align_up(p->sa_ws_bytes, A));
  // Activations: one arena per lane. Ops of a lane run in list order on one stream, so recycling a tensor for a later
  // tensor of the same lane is safe; tensors that another lane touches are never recycled.
  std::vector<size_t> sz(h->tensors.size(), 0);
  std::vector<char> live(h->tensors.size(), 0);
  std::vector<FreeList> arena(h->n_lanes);
  for (size_t i = 0; i < h->ops.size(); ++i) {
    for (size_t t = 0; t < h->tensors.size(); ++t) {
      const TensorInfo& ti = h->tensors[t];
      if (ti.kind == T_WS && ti.first_def == (int)i && !live[t]) {
        sz[t] = align_up((size_t)ti.elems * B * ti.esize, A);
        p->off[t] = arena[ti.home_lane].alloc(sz[t]);
        live[t] = 1;
      }
    }
    for (size_t t = 0; t < h->tensors.size(); ++t) {
      const TensorInfo& ti = h->tensors[t];
      if (ti.kind == T_WS && live[t] && !ti.keep && !ti.cross && ti.last_use == (int)i) { arena[ti.home_lane].release(p->off[t], sz[t]); live[t] = 2; }
    }
  }
  std::vector<size_t> lane_base(h->n_lanes, 0);
  for (int l = 0; l < h->n_lanes; ++l) { lane_base[l] = fl.top; fl.top += align_up(arena[l].top, A); }
  for (size_t t = 0; t < h->tensors.size(); ++t)
    if (h->tensors[t].kind == T_WS) p->off[t] += lane_base[h->tensors[t].home_lane];
  p->ws_bytes = fl.top;
  void* ws = nullptr;
  if (cudaMalloc(&ws, p->ws_bytes) != cudaSuccess) {
    cudaGetLastError();
    return fail(HRP_ERR_NOMEM, "cudaMalloc of %zu workspace bytes for batch %d failed", p->ws_bytes, B);
  }
  p->ws = static_cast<char*>(ws);
  HRP_CUDA(cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming));
  HRP_CUDA(cudaMemset(p->ws + p->io_flags, 0, 8));
  p->flags_state = 0;
  p->stamp = ++h->clock;
  *out = p.get();
  h->plans[B * 8 + slot] = std::move(p);
  return HRP_OK;
}

struct IoPtrs {
  const float *x_reg, *x_root, *k_value, *Kmat;
  float* out;
  const float *init_pose = nullptr, *init_rot = nullptr;   // optional per-frame initial states (full_net.py:268-272)
  const int* flags = nullptr;                               // graph path: which of the two staged overrides are live
  const uint8_t *x_reg_u8 = nullptr, *x_root_u8 = nullptr;  // uint8 NCHW crops instead of x_reg / x_root (hrp_forward_u8): `/ 255.` on device
};

// which ops run_ops enqueues: the ops that read the caller's images (stem / stem_pack, always the first of their lane)
// run AHEAD of the graph on the caller's pointers, so the graph holds no image staging and no 50 MB copy precedes it
enum Part { PART_ALL = 0, PART_PRE = 1, PART_GRAPH = 2 };

struct Profile {
  float* ms; int64_t* launches; double* flops;
  cudaEvent_t e0, e1;
  FILE* dump = nullptr;    // HRP_DUMP_OPS=<path>: one CSV line per op (development / profiles/)
};

// `lanes`: only while capturing the graph. Ops are enqueued on their lane's stream (lane 0 = `st`), cross-lane producers
// signal through events, and every lane joins `st` at the end, so the instantiated graph carries the true dependency DAG
// of the network instead of a serial chain. Without it (eager / profiling) everything runs in list order on `st`.
int run_ops(hrp_handle* h, Plan* p, const IoPtrs& io, cudaStream_t st, Profile* prof, bool lanes = false, Part part = PART_ALL,
            cudaEvent_t root_done = nullptr, int64_t* launched = nullptr) {
  const int B = p->B;
  int64_t offs[HRP_NUM_FIELDS + 1];
  record_floats(h, B, offs);
  auto ptr = [&](int id) -> void* {
    if (id < 0) return nullptr;
    const TensorInfo& t = h->tensors[id];
    switch (t.kind) {
      case T_WS: return p->ws + p->off[id];
      case T_XREG: return const_cast<float*>(io.x_reg);
      case T_XROOT: return const_cast<float*>(io.x_root);
      case T_KVAL: return const_cast<float*>(io.k_value);
      case T_KMAT: return const_cast<float*>(io.Kmat);
      case T_FIELD: return io.out + offs[t.field];
      case T_CONST: return const_cast<void*>(t.cptr);
      case T_INITP: return const_cast<float*>(io.init_pose);
      case T_INITR: return const_cast<float*>(io.init_rot);
      case T_FLAGS: return const_cast<int*>(io.flags);
    }
    return nullptr;
  };
  auto reads_image = [&](const OpDesc& o) { return o.in == h->t_xreg || o.in == h->t_xroot; };
  int64_t launches = 0;
  const int f16 = h->cfg.precision == HRP_PREC_F16 ? 1 : 0;
  const int bf16 = h->cfg.precision == HRP_PREC_BF16 ? 1 : (f16 ? 2 : 0);       // activation element type: 0 fp32, 1 bf16, 2 IEEE half
  const int tf32 = (h->cfg.precision == HRP_PREC_TF32 || h->cfg.precision == HRP_PREC_TF32X3) ? 1 : 0;
  std::vector<cudaEvent_t> op_event(h->ops.size(), nullptr);
  auto new_event = [&]() -> cudaEvent_t {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    h->events.push_back(e);
    return e;
  };
  auto stream_of = [&](int lane) { return (lanes && lane > 0) ? h->lane_stream[lane] : st; };
  if (lanes) {                                   // fork: every lane starts after whatever precedes the graph on `st`
    cudaEvent_t e = new_event();
    if (!e) return fail(HRP_ERR_CUDA, "cudaEventCreate failed");
    HRP_CUDA(cudaEventRecord(e, st));
    for (int l = 1; l < h->n_lanes; ++l) HRP_CUDA(cudaStreamWaitEvent(h->lane_stream[l], e, 0));
  }
  for (size_t oi = 0; oi < h->ops.size(); ++oi) {
    const OpDesc& o = h->ops[oi];
    if (part != PART_ALL && reads_image(o) != (part == PART_PRE)) continue;
    cudaStream_t st_op = stream_of(o.lane);
    if (lanes)
      for (int j : h->waits[oi]) if (op_event[j]) HRP_CUDA(cudaStreamWaitEvent(st_op, op_event[j], 0));
    if (prof) HRP_CUDA(cudaEventRecord(prof->e0, st));
    if (lanes && h->timeline) stamp_kernel<<<1, 1, 0, st_op>>>(h->timeline + 2 * oi);
    int n_launch = 1;
    switch (o.kind) {
      case OP_STEM: {
        const Layer& L = h->layers[o.layer];
        const float* x = static_cast<const float*>(ptr(o.in));
        if (io.x_reg_u8) {                                   // uint8 crops: / 255. into the plan's fp32 staging first
          const bool root = o.in == h->t_xroot;
          float* dst = reinterpret_cast<float*>(p->ws + p->io_xf32) + (root ? (size_t)B * 3 * 256 * 256 : 0);
          HRP_TRY(u8_to_f32_launch(root ? io.x_root_u8 : io.x_reg_u8, dst, (size_t)B * 3 * 256 * 256, st_op));
          x = dst; ++n_launch;
        }
        HRP_TRY(stem_conv_launch(x, L.w, L.bias, ptr(o.out), B, o.Hi, o.Wi, o.Ho, o.Wo, o.KH, o.KW, o.pad_h, bf16 ? 1 : (tf32 ? 2 : 0), st_op));
        break;
      }
      case OP_CONV: {
        const Layer& L = h->layers[o.layer];
        ConvArgs a{};
        a.in = ptr(o.in); a.w = o.cls == CLS_CONV_TC ? L.w_tc : static_cast<const void*>(L.w); a.bias = L.bias; a.res = ptr(o.res); a.out = ptr(o.out);
        a.B = B; a.Hi = o.Hi; a.Wi = o.Wi; a.Cin = o.Cin; a.Ho = o.Ho; a.Wo = o.Wo; a.Cout = o.Cout;
        a.KH = o.KH; a.KW = o.KW; a.stride = o.stride; a.pad_h = o.pad_h; a.pad_w = o.pad_w;
        a.out_sy = o.out_sy; a.out_sx = o.out_sx; a.out_oy = o.out_oy; a.out_ox = o.out_ox; a.Ho_full = o.Ho_full; a.Wo_full = o.Wo_full;
        a.relu = o.relu; a.out_nchw = o.out_nchw; a.ld_out = o.ld; a.out_coff = 0; a.res_after_act = o.res_after_act; a.x3 = o.x3; a.f16 = f16;
        if (o.stem_tc) {
          // view of the packed image: {window 32 el, ox (2-pixel stride), padded row, frame}; the window of output
          // column ox starts at padded x = 2*ox + STEM_PAD - pad (o.pad_w carries the real padding)
          const size_t es = tf32 ? 4 : 2, px = 4 * es, row = (size_t)STEM_WP * px;
          a.pad_w = 0;
          a.in = static_cast<const char*>(a.in) + (size_t)(STEM_PAD - o.pad_w) * px;
          a.tma_custom = 1;
          a.tm_gdim[0] = 32; a.tm_gdim[1] = 128; a.tm_gdim[2] = STEM_HP; a.tm_gdim[3] = (unsigned long long)B;
          a.tm_gstr[0] = 2 * px; a.tm_gstr[1] = row; a.tm_gstr[2] = (size_t)STEM_HP * row;
          a.tm_box[0] = 32; a.tm_box[1] = 128; a.tm_box[2] = 1; a.tm_box[3] = 1;
        }
        if (lanes) a.grid_pct = (h->lane_pct_auto && B < 32) ? 2 * h->lane_pct[o.lane] : h->lane_pct[o.lane];
        if (o.out_nchw && h->sa_fused) a.sa_partial = reinterpret_cast<float*>(p->ws + p->sa_ws);
        if (o.cls == CLS_CONV_TC) HRP_TRY(conv_tc_launch(a, tf32, tf32 && !o.out_nchw, st_op));
        else HRP_TRY(conv_f32_launch(a, st_op));
        break;
      }
      case OP_BLOCK: {
        const Layer &L1 = h->layers[o.layer], &L2 = h->layers[o.layer2];
        BlockArgs a{};
        a.x = ptr(o.in); a.w1 = L1.w_tc; a.w2 = L2.w_tc; a.b1 = L1.bias; a.b2 = L2.bias; a.out = ptr(o.out);
        a.B = B; a.H = o.Hi; a.W = o.Wi; a.C = o.Cin; a.f16 = f16;
        if (lanes) a.grid_pct = (h->lane_pct_auto && B < 32) ? 2 * h->lane_pct[o.lane] : h->lane_pct[o.lane];
        HRP_TRY(conv_block_launch(a, st_op));
        break;
      }
      case OP_CHAIN: {
        ChainArgs a{};
        a.x = ptr(o.in); a.out = ptr(o.out); a.nconv = o.n_chain;
        for (int k = 0; k < o.n_chain; ++k) { a.w[k] = h->layers[o.chain[k]].w_tc; a.b[k] = h->layers[o.chain[k]].bias; }
        a.B = B; a.H = o.Hi; a.W = o.Wi; a.C = o.Cin; a.f16 = f16;
        static const int min_b = [] { const char* v = getenv("HRP_CHAIN_MIN_B"); return v ? atoi(v) : 8; }();
        if (B >= min_b) {
          if (conv_chain_supported(a)) HRP_TRY(conv_chain_launch(a, st_op));
          else HRP_TRY(conv_roll_launch(a, ptr(o.out2), ptr(o.out3), st_op));     // 64 channels: in-place rolling buffer (conv_roll.cu)
          break;
        }
        // small batch: the same eight convs as separate launches, x -> T -> U -> T -> V -> T -> U -> T -> out
        void* T = ptr(o.out2);
        void* blk[5] = {ptr(o.in), ptr(o.out3), ptr(o.out4), ptr(o.out3), ptr(o.out)};
        for (int k = 0; k < o.n_chain; ++k) {
          ConvArgs c{};
          const bool second = (k & 1) != 0;
          c.in = second ? T : blk[k / 2]; c.w = a.w[k]; c.bias = a.b[k]; c.res = second ? blk[k / 2] : nullptr; c.out = second ? blk[k / 2 + 1] : T;
          c.B = B; c.Hi = c.Ho = c.Ho_full = o.Hi; c.Wi = c.Wo = c.Wo_full = o.Wi; c.Cin = c.Cout = c.ld_out = o.Cin;
          c.KH = c.KW = 3; c.stride = 1; c.pad_h = c.pad_w = 1; c.out_sy = c.out_sx = 1; c.relu = 1; c.f16 = f16;
          if (lanes) c.grid_pct = (h->lane_pct_auto && B < 32) ? 2 * h->lane_pct[o.lane] : h->lane_pct[o.lane];
          HRP_TRY(conv_tc_launch(c, 0, 0, st_op));
        }
        n_launch = o.n_chain;
        break;
      }
      case OP_STEM_PACK:
        if (io.x_reg_u8) HRP_TRY(stem_pack_u8_launch(o.in == h->t_xroot ? io.x_root_u8 : io.x_reg_u8, ptr(o.out), B, tf32 ? (o.x3 ? 2 : 1) : (f16 ? 3 : 0), st_op));
        else HRP_TRY(stem_pack_launch(static_cast<const float*>(ptr(o.in)), ptr(o.out), B, tf32 ? (o.x3 ? 2 : 1) : (f16 ? 3 : 0), st_op));
        break;
      case OP_MAXPOOL:
        HRP_TRY(maxpool3x3s2_launch(ptr(o.in), ptr(o.out), B, o.Hi, o.Wi, o.Cin, bf16, st_op));
        break;
      case OP_FUSE: {
        FuseArgs a{};
        for (int k = 0; k < o.n_same; ++k) a.same[k] = ptr(o.same[k]);
        for (int k = 0; k < o.n_low; ++k) { a.low[k] = ptr(o.low[k]); a.shift[k] = o.shift[k]; }
        a.n_same = o.n_same; a.n_low = o.n_low; a.out = ptr(o.out); a.B = B; a.H = o.Ho; a.W = o.Wo; a.C = o.Cout; a.relu = o.relu; a.round_tf32 = tf32 && !o.x3;
        HRP_TRY(fuse_sum_launch(a, bf16, st_op));
        break;
      }
      case OP_AVGPOOL:
        HRP_TRY(avgpool_launch(ptr(o.in), static_cast<float*>(ptr(o.out)), B, o.Hi * o.Wi, o.Cin, bf16, st_op));
        break;
      case OP_DEPTH:
        if (o.relu)          // add_fc / multi_kp variants
          HRP_TRY(depth_head_ex_launch(static_cast<const float*>(ptr(o.in)), static_cast<const float*>(ptr(o.same[0])), static_cast<const float*>(ptr(o.same[1])),
                                       static_cast<const float*>(ptr(o.same[2])), static_cast<const float*>(ptr(o.same[3])), o.bptr,
                                       static_cast<const float*>(ptr(o.in2)), static_cast<float*>(ptr(o.out)), o.out2 >= 0 ? static_cast<float*>(ptr(o.out2)) : nullptr,
                                       B, o.Cin, o.Cout, o.N, o.dof, st_op));
        else
          HRP_TRY(depth_head_launch(static_cast<const float*>(ptr(o.in)), o.wptr, o.bptr, static_cast<const float*>(ptr(o.in2)), static_cast<float*>(ptr(o.out)), B, o.Cin, st_op));
        break;
      case OP_HEADS:
        HRP_TRY(heads_affine_launch(static_cast<const float*>(ptr(o.in)), static_cast<const float*>(ptr(o.same[0])), static_cast<const float*>(ptr(o.same[1])),
                                    static_cast<const float*>(ptr(o.same[2])), static_cast<const float*>(ptr(o.same[3])), static_cast<const float*>(ptr(o.in2)),
                                    static_cast<const float*>(ptr(o.in3)), static_cast<const int*>(ptr(o.in4)), static_cast<float*>(ptr(o.out3)),
                                    static_cast<float*>(ptr(o.out)), static_cast<float*>(ptr(o.out2)), B, o.Cin, o.dof, o.N, o.relu, st_op));
        break;
      case OP_JOINTMAP:
        HRP_TRY(joint_map_head_launch(ptr(o.in), o.wptr, o.bptr, static_cast<const float*>(ptr(o.same[0])), static_cast<float*>(ptr(o.out)), B, o.Hi * o.Wi, o.Cin,
                                      o.dof, bf16, st_op));
        break;
      case OP_SOFTARGMAX: {
        if (h->sa_fused) {       // the heatmap conv already wrote 32 partial states per (frame, keypoint)
          HRP_TRY(softargmax_finalize_launch(reinterpret_cast<const float*>(p->ws + p->sa_ws), B, h->nkpt, 64 * 64 / 128, 64, 64, 64,
                                             static_cast<const float*>(ptr(o.in2)), static_cast<const float*>(ptr(o.in3)), h->cfg.depth_factor,
                                             h->cfg.image_size, h->ref_kp, h->cfg.fix_root, static_cast<float*>(ptr(o.out)), static_cast<float*>(ptr(o.out2)),
                                             static_cast<float*>(ptr(o.out3)), static_cast<float*>(ptr(o.out4)), static_cast<float*>(ptr(o.out5)), st_op));
          break;
        }
        int nl = 0;
        HRP_TRY(softargmax_launch(static_cast<const float*>(ptr(o.in)), B, h->nkpt, 64, 64, 64, static_cast<const float*>(ptr(o.in2)), static_cast<const float*>(ptr(o.in3)),
                                  h->cfg.depth_factor, h->cfg.image_size, h->ref_kp, h->cfg.fix_root, static_cast<float*>(ptr(o.out)), static_cast<float*>(ptr(o.out2)),
                                  p->ws + p->sa_ws, p->sa_ws_bytes, static_cast<float*>(ptr(o.out3)), static_cast<float*>(ptr(o.out4)), static_cast<float*>(ptr(o.out5)), st_op, &nl));
        n_launch = nl;
        break;
      }
      case OP_FK:
        HRP_TRY(fk_launch(h->fk, static_cast<const float*>(ptr(o.in)), static_cast<const float*>(ptr(o.in2)), static_cast<const float*>(ptr(o.in3)), static_cast<const float*>(ptr(o.in4)), B,
                          static_cast<float*>(ptr(o.out)), static_cast<float*>(ptr(o.out2)), st_op));
        break;
    }
    launches += n_launch;
    if (root_done && o.kind == OP_DEPTH) HRP_CUDA(cudaEventRecord(root_done, st));   // DepthNet done (full_net.py:337-340)
    if (lanes && h->timeline) stamp_kernel<<<1, 1, 0, st_op>>>(h->timeline + 2 * oi + 1);
    if (lanes && h->signals[oi]) {
      op_event[oi] = new_event();
      if (!op_event[oi]) return fail(HRP_ERR_CUDA, "cudaEventCreate failed");
      HRP_CUDA(cudaEventRecord(op_event[oi], st_op));
    }
    if (prof) {
      HRP_CUDA(cudaEventRecord(prof->e1, st));
      HRP_CUDA(cudaEventSynchronize(prof->e1));
      float ms = 0.f;
      HRP_CUDA(cudaEventElapsedTime(&ms, prof->e0, prof->e1));
      prof->ms[o.cls] += ms; prof->launches[o.cls] += n_launch; prof->flops[o.cls] += o.flops * B;
      if (prof->dump)
        fprintf(prof->dump, "%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%.4f,%.3f,%d\n", (int)o.kind, o.cls, B, o.Hi, o.Wi, o.Cin, o.Ho, o.Wo, o.Cout, o.KH,
                o.stride, o.res >= 0 ? 1 : 0, o.out_nchw, ms, ms > 0 ? o.flops * B / (ms * 1e-3) / 1e12 : 0.0, o.lane);
    }
  }
  if (lanes)                                     // join
    for (int l = 1; l < h->n_lanes; ++l) {
      cudaEvent_t e = new_event();
      if (!e) return fail(HRP_ERR_CUDA, "cudaEventCreate failed");
      HRP_CUDA(cudaEventRecord(e, h->lane_stream[l]));
      HRP_CUDA(cudaStreamWaitEvent(st, e, 0));
    }
  if (launched) *launched = launches; else h->last_launches = launches;
  return HRP_OK;
}

int check_forward_args(hrp_handle* h, const float* x_reg, const float* x_root, const float* k_value, const float* Kmat, int B, float* out) {
  if (!h) return fail(HRP_ERR_INVALID, "hrp_forward: null handle");
  if (!h->finalized) return fail(HRP_ERR_STATE, "hrp_forward: weights not finalized");
  if (B <= 0 || B > 4096) return fail(HRP_ERR_INVALID, "hrp_forward: batch %d out of range [1, 4096]", B);
  if (!x_reg || !x_root || !k_value || !Kmat || !out) return fail(HRP_ERR_INVALID, "hrp_forward: null pointer");
  return HRP_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int hrp_create(const hrp_config* cfg, const hrp_fk_program* robot, int device, hrp_handle** out) {
  if (!cfg || !robot || !out) return fail(HRP_ERR_INVALID, "hrp_create: null argument");
  if (cfg->backbone != HRP_BACKBONE_RESNET50 && cfg->backbone != HRP_BACKBONE_HRNET32)
    return fail(HRP_ERR_INVALID, "hrp_create: unsupported backbone %d (resnet50 or hrnet32)", cfg->backbone);
  if (cfg->precision < HRP_PREC_FP32 || cfg->precision > HRP_PREC_F16)
    return fail(HRP_ERR_INVALID, "hrp_create: unknown precision %d (0 fp32, 1 tf32, 2 bf16, 3 tf32x3, 4 f16)", cfg->precision);
  if (cfg->n_iter < 1 || cfg->n_iter > 16) return fail(HRP_ERR_INVALID, "hrp_create: n_iter %d out of range", cfg->n_iter);
  if (cfg->image_size != 256.0f) return fail(HRP_ERR_INVALID, "hrp_create: only 256x256 inputs are supported (got %g)", cfg->image_size);
  if (cfg->reg_joint_map) {
    if (cfg->backbone != HRP_BACKBONE_RESNET50) return fail(HRP_ERR_INVALID, "hrp_create: reg_joint_map needs the ResNet keypoint backbone (full_net.py:377 reads x_out)");
    if (robot->dof > 16) return fail(HRP_ERR_INVALID, "hrp_create: reg_joint_map supports up to 16 joints");
    for (int i = 0; i < 3; ++i)
      if (cfg->joint_conv_dim[i] <= 0 || cfg->joint_conv_dim[i] % 32) return fail(HRP_ERR_INVALID, "hrp_create: joint_conv_dim[%d] = %d must be a positive multiple of 32", i, cfg->joint_conv_dim[i]);
  }
  if (cfg->depth_num < 0 || cfg->depth_num > HRP_FK_MAX_KP || cfg->depth_root < 0 || cfg->depth_root >= std::max(1, cfg->depth_num))
    return fail(HRP_ERR_INVALID, "hrp_create: depth_num %d / depth_root %d out of range", cfg->depth_num, cfg->depth_root);
  std::unique_ptr<hrp_handle> h(new hrp_handle());
  h->cfg = *cfg;
  h->device = device;
  HRP_TRY(hrp_fk_create(robot, &h->fk));
  h->dof = robot->dof; h->nkpt = robot->nkpt; h->ref_kp = robot->root_kp;
  const int fw[HRP_NUM_FIELDS] = {h->dof, 6, 3, 2, 1, h->nkpt * 3, h->nkpt * 3, h->nkpt * 3, h->nkpt * 2, h->nkpt * 2, cfg->depth_num};
  for (int f = 0; f < HRP_NUM_FIELDS; ++f) h->field_width[f] = fw[f];
  build_specs(h.get());
  *out = h.release();
  return HRP_OK;
}

extern "C" void hrp_destroy(hrp_handle* h) {
  if (!h) return;
  if (h->timeline) {                                 // stamps of the most recent graph replay
    cudaDeviceSynchronize();
    std::vector<unsigned long long> t(h->ops.size() * 2);
    cudaMemcpy(t.data(), h->timeline, t.size() * 8, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(getenv("HRP_TIMELINE"), "w")) {
      fprintf(f, "op,kind,lane,Hi,Cin,Cout,k,stride,start_ns,end_ns\n");
      unsigned long long t0 = ~0ull;
      for (auto v : t) if (v && v < t0) t0 = v;
      for (size_t i = 0; i < h->ops.size(); ++i) {
        const OpDesc& o = h->ops[i];
        fprintf(f, "%zu,%d,%d,%d,%d,%d,%d,%d,%llu,%llu\n", i, (int)o.kind, o.lane, o.Hi, o.Cin, o.Cout, o.KH, o.stride, t[2 * i] - t0, t[2 * i + 1] - t0);
      }
      fclose(f);
    }
    cudaFree(h->timeline);
  }
  for (auto& kv : h->plans) destroy_plan(kv.second.get());
  for (void* p : h->dev_allocs) cudaFree(p);
  if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
  for (int l = 1; l < kMaxLanes; ++l) if (h->lane_stream[l]) cudaStreamDestroy(h->lane_stream[l]);
  for (cudaEvent_t e : h->events) cudaEventDestroy(e);
  hrp_fk_destroy(h->fk);
  delete h;
}

extern "C" int hrp_num_weights(const hrp_handle* h) { return h ? (int)h->specs.size() : 0; }
extern "C" const char* hrp_weight_name(const hrp_handle* h, int i) {
  return (h && i >= 0 && i < (int)h->specs.size()) ? h->specs[i].name.c_str() : nullptr;
}
extern "C" int hrp_weight_shape(const hrp_handle* h, int i, int64_t* shape, int* ndim) {
  if (!h || i < 0 || i >= (int)h->specs.size() || !shape || !ndim) return fail(HRP_ERR_INVALID, "hrp_weight_shape: bad argument");
  *ndim = (int)h->specs[i].shape.size();
  for (int k = 0; k < *ndim; ++k) shape[k] = h->specs[i].shape[k];
  return HRP_OK;
}

extern "C" int hrp_set_weight(hrp_handle* h, const char* name, const void* host_ptr, const int64_t* shape, int ndim, int dtype) {
  if (!h || !name) return fail(HRP_ERR_INVALID, "hrp_set_weight: null argument");
  if (h->finalized) return fail(HRP_ERR_STATE, "hrp_set_weight: weights already finalized");
  auto it = h->spec_index.find(name);
  if (it == h->spec_index.end()) return fail(HRP_ERR_WEIGHT, "hrp_set_weight: tensor '%s' is not part of this network", name);
  const WeightSpec& sp = h->specs[it->second];
  if (sp.optional) return HRP_OK;  // num_batches_tracked
  if (dtype != HRP_F32) return fail(HRP_ERR_WEIGHT, "hrp_set_weight: '%s' must be float32", name);
  if (!host_ptr) return fail(HRP_ERR_INVALID, "hrp_set_weight: null data for '%s'", name);
  if (ndim != (int)sp.shape.size()) return fail(HRP_ERR_WEIGHT, "hrp_set_weight: '%s' has %d dims, expected %zu", name, ndim, sp.shape.size());
  int64_t n = 1;
  for (int k = 0; k < ndim; ++k) {
    if (shape[k] != sp.shape[k]) return fail(HRP_ERR_WEIGHT, "hrp_set_weight: '%s' dim %d is %lld, expected %lld", name, k, (long long)shape[k], (long long)sp.shape[k]);
    n *= shape[k];
  }
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  t.data.assign(static_cast<const float*>(host_ptr), static_cast<const float*>(host_ptr) + n);
  h->host[name] = std::move(t);
  return HRP_OK;
}

extern "C" int hrp_finalize_weights(hrp_handle* h) {
  if (!h) return fail(HRP_ERR_INVALID, "hrp_finalize_weights: null handle");
  if (h->finalized) return fail(HRP_ERR_STATE, "hrp_finalize_weights: already finalized");
  for (const WeightSpec& sp : h->specs)
    if (!sp.optional && !h->host.count(sp.name)) return fail(HRP_ERR_WEIGHT, "hrp_finalize_weights: tensor '%s' was never set", sp.name.c_str());
  HRP_ON_DEVICE(h);
  GraphBuilder gb{h};
  const int st = gb.build();
  if (st != HRP_OK) return st;
  HRP_CUDA(cudaDeviceSynchronize());
  h->host.clear();
  HRP_CUDA(cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
  if (h->n_lanes > kMaxLanes) return fail(HRP_ERR_INVALID, "internal: %d lanes", h->n_lanes);
  for (int l = 1; l < h->n_lanes; ++l) HRP_CUDA(cudaStreamCreateWithFlags(&h->lane_stream[l], cudaStreamNonBlocking));
  if (const char* e = getenv("HRP_NO_LANES")) h->use_lanes = atoi(e) == 0;
  if (const char* e = getenv("HRP_SLOTS")) h->slots = std::max(1, std::min(8, atoi(e)));
  if (getenv("HRP_TIMELINE")) {
    HRP_CUDA(cudaMalloc(&h->timeline, h->ops.size() * 16));
    HRP_CUDA(cudaMemset(h->timeline, 0, h->ops.size() * 16));
  }
  {
    // lanes 0-3 / 4-7: HRNet branches (full resolution first); lane 4 with the ResNet-50 keypoint backbone: its trunk
    const bool two_hrnets = h->cfg.backbone == HRP_BACKBONE_HRNET32;
    const char* e0 = getenv("HRP_PCT_HI");  const int hi = e0 ? atoi(e0) : 25;     // full-resolution branch, ResNet trunk
    const char* e1 = getenv("HRP_PCT_LO");  const int lo = e1 ? atoi(e1) : 25;     // lower-resolution branches
    if (e0 || e1) h->lane_pct_auto = false;
    for (int l = 0; l < kMaxLanes; ++l) h->lane_pct[l] = ((l & 3) == 0 || (!two_hrnets && l == 4)) ? hi : lo;
  }
  h->finalized = true;
  return HRP_OK;
}

extern "C" int hrp_output_offsets(const hrp_handle* h, int B, int64_t* offsets) {
  if (!h || !offsets || B <= 0) return fail(HRP_ERR_INVALID, "hrp_output_offsets: bad argument");
  record_floats(h, B, offsets);
  return HRP_OK;
}

extern "C" size_t hrp_workspace_bytes(hrp_handle* h, int B) {
  if (!h || !h->finalized || B <= 0) return 0;
  DeviceGuard guard(h->device);
  if (guard.err != cudaSuccess) return 0;
  Plan* p = nullptr;
  if (make_plan(h, B, &p) != HRP_OK) return 0;
  return p->ws_bytes;
}

extern "C" int hrp_set_option(hrp_handle* h, const char* name, int64_t value) {
  if (!h || !name) return fail(HRP_ERR_INVALID, "hrp_set_option: null argument");
  if (std::strcmp(name, "cuda_graph") == 0) { h->use_graph = value != 0; return HRP_OK; }
  if (std::strcmp(name, "slots") == 0) {         // plans per batch size, 1..8 (default 4)
    if (value < 1 || value > 8) return fail(HRP_ERR_INVALID, "hrp_set_option: slots must be 1..8");
    h->slots = (int)value; h->next_slot = 0;
    return HRP_OK;
  }
  if (std::strcmp(name, "lane_share_pct") == 0) { // share of the CTA slots one conv launch may take (default 25: tuned for a
    if (value < 5 || value > 100) return fail(HRP_ERR_INVALID, "hrp_set_option: lane_share_pct must be 5..100");   // stream of overlapping
    h->lane_pct_auto = false;
    for (int l = 0; l < kMaxLanes; ++l) h->lane_pct[l] = (int)value;                                               // forwards; 50 gives the
    return HRP_OK;                                                                                                 // lowest single-call latency)
  }
  if (std::strcmp(name, "max_cached_batches") == 0) {   // distinct batch sizes whose plans stay cached (default 4)
    if (value < 1 || value > 64) return fail(HRP_ERR_INVALID, "hrp_set_option: max_cached_batches must be 1..64");
    h->max_batches = (int)value;
    return HRP_OK;
  }
  if (std::strcmp(name, "lanes") == 0) {         // multi-stream graph (default 1); takes effect for graphs not yet captured
    h->use_lanes = value != 0;
    return HRP_OK;
  }
  return fail(HRP_ERR_INVALID, "hrp_set_option: unknown option '%s'", name);
}

extern "C" int64_t hrp_launch_count(const hrp_handle* h) { return h ? h->last_launches : 0; }

namespace {
int forward_impl(hrp_handle* h, const float* x_reg, const float* x_root, const float* k_value, const float* Kmat,
                 const float* init_pose, const float* init_rot, int B, float* out, cudaStream_t st, float* ms3,
                 const uint8_t* x_reg_u8 = nullptr, const uint8_t* x_root_u8 = nullptr) {
  if (x_reg_u8) {                  // the float pointers only serve the null check below
    x_reg = reinterpret_cast<const float*>(x_reg_u8);
    x_root = reinterpret_cast<const float*>(x_root_u8);
  }
  HRP_TRY(check_forward_args(h, x_reg, x_root, k_value, Kmat, B, out));
  if (x_reg_u8) { x_reg = nullptr; x_root = nullptr; }
  HRP_ON_DEVICE(h);
  Plan* p = nullptr;
  const bool graph = h->use_graph && ms3 == nullptr;
  int slot = 0;
  if (graph) {
    // a forward on the same stream as its predecessor cannot overlap it anyway: it reuses that plan, so a caller with one
    // stream only ever pays for one workspace; a forward on another stream takes the next plan
    if (h->last_B == B && st == h->last_stream) slot = h->last_slot_used;
    else { slot = h->next_slot; h->next_slot = (h->next_slot + 1) % h->slots; }
    h->last_stream = st; h->last_B = B; h->last_slot_used = slot;
  }
  HRP_TRY(make_plan(h, B, &p, slot));
  h->last_slot[B] = slot;
  // whatever stream used this plan last must be done with its workspace before this forward touches it
  if (p->used) HRP_CUDA(cudaStreamWaitEvent(st, p->done, 0));
  if (!graph) {
    IoPtrs io{x_reg, x_root, k_value, Kmat, out, init_pose, init_rot, nullptr, x_reg_u8, x_root_u8};
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    if (ms3) {
      HRP_CUDA(cudaEventCreate(&e0)); HRP_CUDA(cudaEventCreate(&e1)); HRP_CUDA(cudaEventCreate(&e2));
      HRP_CUDA(cudaEventRecord(e0, st));
    }
    int rs = run_ops(h, p, io, st, nullptr, false, PART_ALL, e1);
    if (ms3 && rs == HRP_OK) {
      // the reference synchronises its stream after the DepthNet and at the end of the forward (full_net.py:337-340,
      // 452-457) and reports (time_root, time_other, time_whole); here: device time between events, list order = DepthNet first
      cudaEventRecord(e2, st);
      if (cudaEventSynchronize(e2) != cudaSuccess) rs = fail(HRP_ERR_CUDA, "hrp_forward_timed: %s", cudaGetErrorString(cudaGetLastError()));
      else { cudaEventElapsedTime(&ms3[0], e0, e1); cudaEventElapsedTime(&ms3[1], e1, e2); cudaEventElapsedTime(&ms3[2], e0, e2); }
    }
    if (ms3) { cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); }
    HRP_CUDA(cudaEventRecord(p->done, st));
    p->used = true;
    return rs;
  }
  // graph path: the graph reads / writes the plan's static staging (camera, initial states, output record), so one
  // instantiation serves every call; the images are consumed by the pre-graph ops straight from the caller's buffers
  float* s_kv = reinterpret_cast<float*>(p->ws + p->io_kv);
  float* s_K = reinterpret_cast<float*>(p->ws + p->io_K);
  float* s_out = reinterpret_cast<float*>(p->ws + p->io_out);
  float* s_ip = reinterpret_cast<float*>(p->ws + p->io_initp);
  float* s_ir = reinterpret_cast<float*>(p->ws + p->io_initr);
  int* s_flags = reinterpret_cast<int*>(p->ws + p->io_flags);
  if (!p->exec) {
    cudaGraph_t g = nullptr;
    int64_t n_graph = 0;
    HRP_CUDA(cudaStreamBeginCapture(h->capture_stream, cudaStreamCaptureModeThreadLocal));
    const int rs = run_ops(h, p, IoPtrs{nullptr, nullptr, s_kv, s_K, s_out, s_ip, s_ir, s_flags}, h->capture_stream, nullptr,
                           h->use_lanes && h->n_lanes > 1, PART_GRAPH, nullptr, &n_graph);
    cudaError_t ce = cudaStreamEndCapture(h->capture_stream, &g);
    if (rs != HRP_OK) { if (g) cudaGraphDestroy(g); return rs; }
    if (ce != cudaSuccess) return fail(HRP_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&p->exec, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) return fail(HRP_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ce));
    p->launches = n_graph;
  }
  int64_t n_pre = 0;
  HRP_TRY(run_ops(h, p, IoPtrs{x_reg, x_root, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, x_reg_u8, x_root_u8}, st, nullptr, false, PART_PRE, nullptr, &n_pre));
  HRP_CUDA(cudaMemcpyAsync(s_kv, k_value, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  HRP_CUDA(cudaMemcpyAsync(s_K, Kmat, (size_t)B * 9 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (init_pose) HRP_CUDA(cudaMemcpyAsync(s_ip, init_pose, (size_t)B * h->dof * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (init_rot) HRP_CUDA(cudaMemcpyAsync(s_ir, init_rot, (size_t)B * 6 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const int want = (init_pose ? 1 : 0) | (init_rot ? 2 : 0);
  if (p->flags_state != want) {
    HRP_CUDA(cudaMemsetAsync(s_flags, init_pose ? 1 : 0, 4, st));
    HRP_CUDA(cudaMemsetAsync(s_flags + 1, init_rot ? 1 : 0, 4, st));
    p->flags_state = want;
  }
  HRP_CUDA(cudaGraphLaunch(p->exec, st));
  HRP_CUDA(cudaMemcpyAsync(out, s_out, (size_t)record_floats(h, B, nullptr) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  HRP_CUDA(cudaEventRecord(p->done, st));
  p->used = true;
  h->last_launches = p->launches + n_pre;
  return HRP_OK;
}
}  // namespace

extern "C" int hrp_forward(hrp_handle* h, const float* x_reg, const float* x_root, const float* k_value, const float* Kmat,
                           int B, float* out, void* stream) {
  return forward_impl(h, x_reg, x_root, k_value, Kmat, nullptr, nullptr, B, out, (cudaStream_t)stream, nullptr);
}

extern "C" int hrp_forward_ex(hrp_handle* h, const float* x_reg, const float* x_root, const float* k_value, const float* Kmat,
                              const float* init_pose, const float* init_rot, int B, float* out, void* stream) {
  return forward_impl(h, x_reg, x_root, k_value, Kmat, init_pose, init_rot, B, out, (cudaStream_t)stream, nullptr);
}

extern "C" int hrp_forward_timed(hrp_handle* h, const float* x_reg, const float* x_root, const float* k_value, const float* Kmat,
                                 const float* init_pose, const float* init_rot, int B, float* out, float* ms3, void* stream) {
  if (!ms3) return fail(HRP_ERR_INVALID, "hrp_forward_timed: null timing output");
  return forward_impl(h, x_reg, x_root, k_value, Kmat, init_pose, init_rot, B, out, (cudaStream_t)stream, ms3);
}

extern "C" int hrp_forward_u8(hrp_handle* h, const uint8_t* x_reg, const uint8_t* x_root, const float* k_value, const float* Kmat,
                              int B, float* out, void* stream) {
  if (!x_reg || !x_root) return fail(HRP_ERR_INVALID, "hrp_forward_u8: null pointer");
  return forward_impl(h, nullptr, nullptr, k_value, Kmat, nullptr, nullptr, B, out, (cudaStream_t)stream, nullptr, x_reg, x_root);
}

extern "C" int hrp_release_plans(hrp_handle* h) {
  if (!h) return fail(HRP_ERR_INVALID, "hrp_release_plans: null handle");
  HRP_ON_DEVICE(h);
  for (auto& kv : h->plans) destroy_plan(kv.second.get());
  h->plans.clear();
  h->last_slot.clear();
  h->last_B = 0; h->last_stream = nullptr; h->next_slot = 0;
  return HRP_OK;
}

extern "C" int hrp_forward_profile(hrp_handle* h, const float* x_reg, const float* x_root, const float* k_value,
                                   const float* Kmat, int B, float* out, float* ms_by_class, int64_t* launches_by_class,
                                   double* flops_by_class, void* stream) {
  HRP_TRY(check_forward_args(h, x_reg, x_root, k_value, Kmat, B, out));
  if (!ms_by_class || !launches_by_class || !flops_by_class) return fail(HRP_ERR_INVALID, "hrp_forward_profile: null output");
  HRP_ON_DEVICE(h);
  Plan* p = nullptr;
  HRP_TRY(make_plan(h, B, &p));
  for (int c = 0; c < HRP_NUM_CLASSES; ++c) { ms_by_class[c] = 0.f; launches_by_class[c] = 0; flops_by_class[c] = 0.0; }
  Profile prof{ms_by_class, launches_by_class, flops_by_class, nullptr, nullptr};
  HRP_CUDA(cudaEventCreate(&prof.e0));
  HRP_CUDA(cudaEventCreate(&prof.e1));
  if (const char* path = getenv("HRP_DUMP_OPS")) {
    prof.dump = fopen(path, "w");
    if (prof.dump) fprintf(prof.dump, "kind,cls,B,Hi,Wi,Cin,Ho,Wo,Cout,k,stride,res,nchw,ms,tflops,lane\n");
  }
  const int rs = run_ops(h, p, IoPtrs{x_reg, x_root, k_value, Kmat, out}, (cudaStream_t)stream, &prof);
  if (prof.dump) fclose(prof.dump);
  cudaEventDestroy(prof.e0);
  cudaEventDestroy(prof.e1);
  return rs;
}

extern "C" int hrp_debug_tensor(hrp_handle* h, const char* name, int B, float* dst_device, int64_t* numel, void* stream) {
  if (!h || !name || !numel) return fail(HRP_ERR_INVALID, "hrp_debug_tensor: null argument");
  auto it = h->debug.find(name);
  if (it == h->debug.end()) return fail(HRP_ERR_INVALID, "hrp_debug_tensor: unknown tensor '%s' (xf, img_feat, logits, head_iters)", name);
  auto ls = h->last_slot.find(B);
  auto pit = ls == h->last_slot.end() ? h->plans.end() : h->plans.find(B * 8 + ls->second);
  if (pit == h->plans.end()) return fail(HRP_ERR_STATE, "hrp_debug_tensor: no forward has run for batch %d", B);
  const TensorInfo& t = h->tensors[it->second];
  *numel = t.elems * B;
  if (dst_device) HRP_CUDA(cudaMemcpyAsync(dst_device, pit->second->ws + pit->second->off[it->second], (size_t)*numel * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return HRP_OK;
}

// single-layer entry point for parity tests (packs on the fly; synchronises; not a timed path)
extern "C" int hrp_conv2d_nhwc(const float* in, const float* weight_oihw, const float* bias, const float* residual, float* out,
                               int B, int Hi, int Wi, int Cin, int Cout, int KH, int KW, int stride, int pad, int relu,
                               int precision, void* stream) {
  if (!in || !weight_oihw || !out) return fail(HRP_ERR_INVALID, "hrp_conv2d_nhwc: null argument");
  if (precision < HRP_PREC_FP32 || precision > HRP_PREC_F16)
    return fail(HRP_ERR_INVALID, "hrp_conv2d_nhwc: unknown precision %d", precision);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nw = (size_t)Cout * Cin * KH * KW;
  std::vector<float> w(nw), b(Cout, 0.f), wp(nw), bp(Cout);
  HRP_CUDA(cudaMemcpyAsync(w.data(), weight_oihw, nw * 4, cudaMemcpyDeviceToHost, st));
  if (bias) HRP_CUDA(cudaMemcpyAsync(b.data(), bias, (size_t)Cout * 4, cudaMemcpyDeviceToHost, st));
  HRP_CUDA(cudaStreamSynchronize(st));
  pack_conv_f32(w.data(), b.data(), nullptr, nullptr, nullptr, nullptr, Cout, Cin, KH, KW, wp.data(), bp.data());
  ConvArgs a{};
  a.B = B; a.Hi = Hi; a.Wi = Wi; a.Cin = Cin; a.Cout = Cout; a.KH = KH; a.KW = KW; a.stride = stride; a.pad_h = a.pad_w = pad;
  a.Ho = (Hi + 2 * pad - KH) / stride + 1; a.Wo = (Wi + 2 * pad - KW) / stride + 1;
  a.out_sy = a.out_sx = 1; a.Ho_full = a.Ho; a.Wo_full = a.Wo; a.relu = relu; a.ld_out = Cout;
  const size_t n_in = (size_t)B * Hi * Wi * Cin, n_out = (size_t)B * a.Ho * a.Wo * Cout;
  std::vector<void*> tmp;
  auto dalloc = [&](size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(bytes, 16)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    tmp.push_back(d);
    return d;
  };
  int rs = HRP_OK;
  float* db = static_cast<float*>(dalloc((size_t)Cout * 4));
  if (!db) rs = fail(HRP_ERR_NOMEM, "hrp_conv2d_nhwc: out of device memory");
  if (rs == HRP_OK && cudaMemcpyAsync(db, bp.data(), (size_t)Cout * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) rs = fail(HRP_ERR_CUDA, "hrp_conv2d_nhwc: copy failed");
  a.bias = db;
  if (rs == HRP_OK && precision == HRP_PREC_FP32) {
    float* dw = static_cast<float*>(dalloc(nw * 4));
    if (!dw) rs = fail(HRP_ERR_NOMEM, "hrp_conv2d_nhwc: out of device memory");
    if (rs == HRP_OK && cudaMemcpyAsync(dw, wp.data(), nw * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) rs = fail(HRP_ERR_CUDA, "hrp_conv2d_nhwc: copy failed");
    a.in = in; a.w = dw; a.res = residual; a.out = out;
    if (rs == HRP_OK) rs = conv_f32_launch(a, st);
  } else if (rs == HRP_OK) {
    // tensor-core families: operands converted to the family's activation type exactly as a producing layer would
    const int x3 = precision == HRP_PREC_TF32X3;                // 3xTF32: full-fp32 operands, hi/lo split inside the kernel
    const int tf32 = precision == HRP_PREC_TF32 || x3;
    const int f16 = precision == HRP_PREC_F16;
    a.x3 = x3; a.f16 = f16;
    if (!conv_tc_supported(a, tf32)) rs = fail(HRP_ERR_INVALID, "hrp_conv2d_nhwc: shape not supported by the tensor-core family (Cin %% %d, Cout %% 16)", tf32 ? 16 : 32);
    const int rb = conv_tc_row_bytes(a, tf32, nullptr);
    std::vector<uint8_t> img(pack_conv_tc_bytes(KH * KW * Cin, Cout, tf32 ? 1 + x3 : (f16 ? 3 : 0), rb));
    pack_conv_tc(wp.data(), KH * KW * Cin, Cout, tf32 ? 1 + x3 : (f16 ? 3 : 0), rb, img.data());
    void* dw = dalloc(img.size());
    const size_t es = tf32 ? 4 : 2;
    void* din = dalloc(n_in * es);
    void* dres = residual ? dalloc(n_out * es) : nullptr;
    void* dout = tf32 ? static_cast<void*>(out) : dalloc(n_out * es);
    if (rs == HRP_OK && (!dw || !din || !dout || (residual && !dres))) rs = fail(HRP_ERR_NOMEM, "hrp_conv2d_nhwc: out of device memory");
    if (rs == HRP_OK && cudaMemcpyAsync(dw, img.data(), img.size(), cudaMemcpyHostToDevice, st) != cudaSuccess) rs = fail(HRP_ERR_CUDA, "hrp_conv2d_nhwc: copy failed");
    if (rs == HRP_OK && x3 && cudaMemcpyAsync(din, in, n_in * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rs = fail(HRP_ERR_CUDA, "hrp_conv2d_nhwc: copy failed");
    if (rs == HRP_OK && x3 && residual && cudaMemcpyAsync(dres, residual, n_out * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rs = fail(HRP_ERR_CUDA, "hrp_conv2d_nhwc: copy failed");
    if (rs == HRP_OK && !x3) rs = tf32 ? round_tf32_launch(in, static_cast<float*>(din), n_in, st) : cast_f32_to_bf16_launch(in, din, n_in, st, f16);
    if (rs == HRP_OK && !x3 && residual) rs = tf32 ? round_tf32_launch(residual, static_cast<float*>(dres), n_out, st) : cast_f32_to_bf16_launch(residual, dres, n_out, st, f16);
    a.in = din; a.w = dw; a.res = dres; a.out = dout;
    if (rs == HRP_OK) rs = conv_tc_launch(a, tf32, tf32, st);
    if (rs == HRP_OK && !tf32) rs = cast_bf16_to_f32_launch(dout, out, n_out, st, f16);
  }
  if (cudaStreamSynchronize(st) != cudaSuccess && rs == HRP_OK) rs = fail(HRP_ERR_CUDA, "hrp_conv2d_nhwc: %s", cudaGetErrorString(cudaGetLastError()));
  for (void* d : tmp) cudaFree(d);
  return rs;
}

// Micro-benchmark of one tensor-core conv layer in isolation (random operands already in the family's type; the weight
// image is packed once; `iters` back-to-back launches timed with CUDA events on `stream`). Tuning aid for the kernels.
extern "C" int hrp_conv_bench(int precision, int B, int H, int W, int Cin, int Cout, int k, int stride, int with_residual,
                              int iters, float* ms_per_launch, void* stream) {
  if (precision != HRP_PREC_TF32 && precision != HRP_PREC_BF16 && precision != HRP_PREC_F16) return fail(HRP_ERR_INVALID, "hrp_conv_bench: tensor-core families only");
  const int f16 = precision == HRP_PREC_F16;
  if (!ms_per_launch || iters <= 0) return fail(HRP_ERR_INVALID, "hrp_conv_bench: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int tf32 = precision == HRP_PREC_TF32;
  const size_t es = tf32 ? 4 : 2;
  ConvArgs a{};
  a.B = B; a.Hi = H; a.Wi = W; a.Cin = Cin; a.Cout = Cout; a.KH = a.KW = k; a.stride = stride; a.pad_h = a.pad_w = k / 2;
  a.Ho = (H + 2 * a.pad_h - k) / stride + 1; a.Wo = (W + 2 * a.pad_w - k) / stride + 1;
  a.out_sy = a.out_sx = 1; a.Ho_full = a.Ho; a.Wo_full = a.Wo; a.relu = 1; a.ld_out = Cout; a.f16 = f16;
  if (const char* e = getenv("HRP_BENCH_PCT")) a.grid_pct = atoi(e);      // the share of the GPU a launch gets inside the multi-lane graph
  if (!conv_tc_supported(a, tf32)) return fail(HRP_ERR_INVALID, "hrp_conv_bench: shape not supported");
  const size_t n_in = (size_t)B * H * W * Cin, n_out = (size_t)B * a.Ho * a.Wo * Cout, K = (size_t)k * k * Cin;
  std::vector<float> wp(K * Cout), bp(Cout, 0.1f);
  uint32_t lcg = 12345u;
  for (auto& v : wp) { lcg = lcg * 1664525u + 1013904223u; v = ((lcg >> 8) * (1.0f / 16777216.0f) - 0.5f) / std::sqrt((float)K); }
  const int rb = conv_tc_row_bytes(a, tf32, nullptr);
  std::vector<uint8_t> img(pack_conv_tc_bytes((int)K, Cout, tf32 ? 1 : (f16 ? 3 : 0), rb));
  pack_conv_tc(wp.data(), (int)K, Cout, tf32 ? 1 : (f16 ? 3 : 0), rb, img.data());
  void *dw = nullptr, *din = nullptr, *dout = nullptr, *dres = nullptr, *db = nullptr;
  float* tmp = nullptr;
  int rs = HRP_OK;
  auto A = [&](void** ptr, size_t bytes) { if (rs == HRP_OK && cudaMalloc(ptr, std::max<size_t>(bytes, 16)) != cudaSuccess) { cudaGetLastError(); rs = fail(HRP_ERR_NOMEM, "hrp_conv_bench: out of device memory"); } };
  A(&dw, img.size()); A(&din, n_in * es); A(&dout, n_out * es); A(&db, (size_t)Cout * 4); A((void**)&tmp, std::max(n_in, n_out) * 4);
  if (with_residual) A(&dres, n_out * es);
  if (rs == HRP_OK) {
    std::vector<float> host(std::max(n_in, n_out));
    for (auto& v : host) { lcg = lcg * 1664525u + 1013904223u; v = (lcg >> 8) * (1.0f / 16777216.0f) - 0.5f; }
    cudaMemcpyAsync(dw, img.data(), img.size(), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(db, bp.data(), (size_t)Cout * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(tmp, host.data(), host.size() * 4, cudaMemcpyHostToDevice, st);
    rs = tf32 ? round_tf32_launch(tmp, static_cast<float*>(din), n_in, st) : cast_f32_to_bf16_launch(tmp, din, n_in, st, f16);
    if (rs == HRP_OK && with_residual) rs = tf32 ? round_tf32_launch(tmp, static_cast<float*>(dres), n_out, st) : cast_f32_to_bf16_launch(tmp, dres, n_out, st, f16);
    cudaStreamSynchronize(st);
  }
  a.in = din; a.w = dw; a.bias = static_cast<const float*>(db); a.res = dres; a.out = dout;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (rs == HRP_OK) {
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // with_residual == 8: time the 8-conv branch chain (conv_chain.cu) on this shape instead, the same weights for every conv
    ChainArgs ch{};
    const bool chain = with_residual == 8;
    if (chain) {
      ch.x = din; ch.out = dout; ch.nconv = 8; ch.B = B; ch.H = H; ch.W = W; ch.C = Cin; ch.f16 = f16;
      for (int j = 0; j < 8; ++j) { ch.w[j] = dw; ch.b[j] = static_cast<const float*>(db); }
      if (k != 3 || stride != 1 || Cin != Cout || tf32 || !(conv_chain_supported(ch) || conv_roll_supported(ch))) rs = fail(HRP_ERR_INVALID, "hrp_conv_bench: shape not taken by the chain kernels");
    }
    void *sc0 = nullptr, *sc1 = nullptr;
    const bool roll = chain && rs == HRP_OK && !conv_chain_supported(ch);
    if (roll) { A(&sc0, n_out * es); A(&sc1, n_out * es); }
    auto launch = [&](cudaStream_t s_) { return chain ? (roll ? conv_roll_launch(ch, sc0, sc1, s_) : conv_chain_launch(ch, s_)) : conv_tc_launch(a, tf32, tf32, s_); };
    for (int i = 0; i < 3 && rs == HRP_OK; ++i) rs = launch(st);
    // the timed launches replay from a CUDA graph so that host-side launch cost (tensor-map encoding, ~15 us) is not what
    // gets measured for kernels shorter than that
    cudaStream_t cs = nullptr;
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    cudaStreamSynchronize(st);
    if (rs == HRP_OK && cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      for (int i = 0; i < iters && rs == HRP_OK; ++i) rs = launch(cs);
      if (cudaStreamEndCapture(cs, &g) != cudaSuccess || cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) rs = fail(HRP_ERR_CUDA, "hrp_conv_bench: graph capture failed");
    }
    if (rs == HRP_OK) cudaGraphLaunch(ge, st);      // warm replay
    cudaEventRecord(e0, st);
    if (rs == HRP_OK) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess && rs == HRP_OK) rs = fail(HRP_ERR_CUDA, "hrp_conv_bench: %s", cudaGetErrorString(cudaGetLastError()));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_per_launch = ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ge) cudaGraphExecDestroy(ge);
    if (g) cudaGraphDestroy(g);
    if (cs) cudaStreamDestroy(cs);
    if (sc0) cudaFree(sc0);
    if (sc1) cudaFree(sc1);
  }
  for (void* d : {dw, din, dout, dres, db, (void*)tmp}) if (d) cudaFree(d);
  return rs;
}

// One BasicBlock through the fused kernel (layer-level parity test): x NHWC [B,H,W,C] fp32 (rounded to bf16 here), OIHW
// fp32 weights and biases of the two convs; out fp32. Synchronises; not a timed path.
extern "C" int hrp_basic_block_nhwc(const float* x, const float* w1_oihw, const float* b1, const float* w2_oihw, const float* b2,
                                    float* out, int B, int H, int W, int C, void* stream) {
  if (!x || !w1_oihw || !w2_oihw || !out) return fail(HRP_ERR_INVALID, "hrp_basic_block_nhwc: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  BlockArgs a{};
  a.B = B; a.H = H; a.W = W; a.C = C;
  if (!conv_block_supported(a)) return fail(HRP_ERR_INVALID, "hrp_basic_block_nhwc: block not supported by the fused kernel (C=%d, %dx%d)", C, H, W);
  const size_t nw = (size_t)C * C * 9, n = (size_t)B * H * W * C;
  std::vector<float> w(nw), b(C), wp(nw), bp(C);
  std::vector<void*> tmp;
  auto dalloc = [&](size_t bytes) -> void* { void* d = nullptr; if (cudaMalloc(&d, std::max<size_t>(bytes, 16)) != cudaSuccess) { cudaGetLastError(); return nullptr; } tmp.push_back(d); return d; };
  ConvArgs g{};
  g.B = 1; g.Hi = g.Ho = H; g.Wi = g.Wo = W; g.Cin = g.Cout = C; g.KH = g.KW = 3; g.stride = 1; g.pad_h = g.pad_w = 1; g.ld_out = C;
  const int rb = conv_tc_row_bytes(g, 0, nullptr);
  int rs = rb == 64 ? HRP_OK : fail(HRP_ERR_INVALID, "hrp_basic_block_nhwc: unexpected operand row width %d", rb);
  void* dw[2] = {nullptr, nullptr};
  float* db[2] = {nullptr, nullptr};
  const float* wsrc[2] = {w1_oihw, w2_oihw};
  const float* bsrc[2] = {b1, b2};
  for (int k = 0; k < 2 && rs == HRP_OK; ++k) {
    cudaMemcpyAsync(w.data(), wsrc[k], nw * 4, cudaMemcpyDeviceToHost, st);
    if (bsrc[k]) cudaMemcpyAsync(b.data(), bsrc[k], (size_t)C * 4, cudaMemcpyDeviceToHost, st); else std::fill(b.begin(), b.end(), 0.f);
    cudaStreamSynchronize(st);
    pack_conv_f32(w.data(), b.data(), nullptr, nullptr, nullptr, nullptr, C, C, 3, 3, wp.data(), bp.data());
    std::vector<uint8_t> img(pack_conv_tc_bytes(9 * C, C, 0, rb));
    pack_conv_tc(wp.data(), 9 * C, C, 0, rb, img.data());
    dw[k] = dalloc(img.size());
    db[k] = static_cast<float*>(dalloc((size_t)C * 4));
    if (!dw[k] || !db[k]) { rs = fail(HRP_ERR_NOMEM, "hrp_basic_block_nhwc: out of device memory"); break; }
    cudaMemcpy(dw[k], img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db[k], bp.data(), (size_t)C * 4, cudaMemcpyHostToDevice);
  }
  void* dx = dalloc(n * 2);
  void* dout = dalloc(n * 2);
  if (rs == HRP_OK && (!dx || !dout)) rs = fail(HRP_ERR_NOMEM, "hrp_basic_block_nhwc: out of device memory");
  if (rs == HRP_OK) rs = cast_f32_to_bf16_launch(x, dx, n, st);
  a.x = dx; a.w1 = dw[0]; a.w2 = dw[1]; a.b1 = db[0]; a.b2 = db[1]; a.out = dout;
  if (rs == HRP_OK) rs = conv_block_launch(a, st);
  if (rs == HRP_OK) rs = cast_bf16_to_f32_launch(dout, out, n, st);
  if (cudaStreamSynchronize(st) != cudaSuccess && rs == HRP_OK) rs = fail(HRP_ERR_CUDA, "hrp_basic_block_nhwc: %s", cudaGetErrorString(cudaGetLastError()));
  for (void* d : tmp) cudaFree(d);
  return rs;
}

// A chain of `nblocks` BasicBlocks through the one-launch branch kernel (conv_chain.cu); layer-level parity tests.
// w_oihw: [2*nblocks][C][C][3][3], b: [2*nblocks][C] (may be null), device fp32; x / out NHWC fp32 (cast to / from bf16 here).
extern "C" int hrp_basic_chain_nhwc(const float* x, const float* w_oihw, const float* b, int nblocks, float* out, int B, int H, int W, int C,
                                    void* stream) {
  if (!x || !w_oihw || !out) return fail(HRP_ERR_INVALID, "hrp_basic_chain_nhwc: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  ChainArgs a{};
  a.B = B; a.H = H; a.W = W; a.C = C; a.nconv = 2 * nblocks;
  const bool roll = nblocks >= 1 && nblocks <= 4 && !conv_chain_supported(a) && conv_roll_supported(a);
  if (nblocks < 1 || nblocks > 4 || !(conv_chain_supported(a) || roll))
    return fail(HRP_ERR_INVALID, "hrp_basic_chain_nhwc: chain not supported by the fused kernels (C=%d, %dx%d, %d blocks)", C, H, W, nblocks);
  const size_t nw = (size_t)C * C * 9, n = (size_t)B * H * W * C;
  std::vector<float> w(nw), bb(C), wp(nw), bp(C);
  std::vector<void*> tmp;
  auto dalloc = [&](size_t bytes) -> void* { void* d = nullptr; if (cudaMalloc(&d, std::max<size_t>(bytes, 16)) != cudaSuccess) { cudaGetLastError(); return nullptr; } tmp.push_back(d); return d; };
  int rs = HRP_OK;
  for (int k = 0; k < a.nconv && rs == HRP_OK; ++k) {
    cudaMemcpyAsync(w.data(), w_oihw + (size_t)k * nw, nw * 4, cudaMemcpyDeviceToHost, st);
    if (b) cudaMemcpyAsync(bb.data(), b + (size_t)k * C, (size_t)C * 4, cudaMemcpyDeviceToHost, st); else std::fill(bb.begin(), bb.end(), 0.f);
    cudaStreamSynchronize(st);
    pack_conv_f32(w.data(), bb.data(), nullptr, nullptr, nullptr, nullptr, C, C, 3, 3, wp.data(), bp.data());
    std::vector<uint8_t> img(pack_conv_tc_bytes(9 * C, C, 0, 128));
    pack_conv_tc(wp.data(), 9 * C, C, 0, 128, img.data());
    void* dw = dalloc(img.size());
    float* db = static_cast<float*>(dalloc((size_t)C * 4));
    if (!dw || !db) { rs = fail(HRP_ERR_NOMEM, "hrp_basic_chain_nhwc: out of device memory"); break; }
    cudaMemcpy(dw, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bp.data(), (size_t)C * 4, cudaMemcpyHostToDevice);
    a.w[k] = dw; a.b[k] = db;
  }
  void* dx = dalloc(n * 2);
  void* dout = dalloc(n * 2);
  if (rs == HRP_OK && (!dx || !dout)) rs = fail(HRP_ERR_NOMEM, "hrp_basic_chain_nhwc: out of device memory");
  if (rs == HRP_OK) rs = cast_f32_to_bf16_launch(x, dx, n, st);
  a.x = dx; a.out = dout;
  void* sc0 = roll ? dalloc(n * 2) : nullptr;
  void* sc1 = roll ? dalloc(n * 2) : nullptr;
  if (rs == HRP_OK && roll && (!sc0 || !sc1)) rs = fail(HRP_ERR_NOMEM, "hrp_basic_chain_nhwc: out of device memory");
  if (rs == HRP_OK) rs = roll ? conv_roll_launch(a, sc0, sc1, st) : conv_chain_launch(a, st);
  if (rs == HRP_OK) rs = cast_bf16_to_f32_launch(dout, out, n, st);
  if (cudaStreamSynchronize(st) != cudaSuccess && rs == HRP_OK) rs = fail(HRP_ERR_CUDA, "hrp_basic_chain_nhwc: %s", cudaGetErrorString(cudaGetLastError()));
  for (void* d : tmp) cudaFree(d);
  return rs;
}
