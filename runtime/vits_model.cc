// See vits_model.h.  Call sequence of ForwardEncoder = SynthesizerTrn.export_encoder_forward (models.py:346-355):
// speaker embedding, text encoder, duration predictor, length regulation, prior expansion + sampling, flow
// inversion; ForwardDecoder = export_decoder_forward (models.py:357-363).
#include "vits_model.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace wetts {

namespace {

const char kMagic[8] = {'W', 'E', 'T', 'T', 'S', 'B', '2', '\0'};

template <typename T>
bool ReadPod(std::ifstream& f, T* v) {
  f.read(reinterpret_cast<char*>(v), sizeof(T));
  return static_cast<bool>(f);
}

struct DeviceBuffer {
  void* p = nullptr;
  explicit DeviceBuffer(size_t bytes) {
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
  }
  ~DeviceBuffer() { cudaFree(p); }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

void CudaCheck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

}  // namespace

bool LoadFlatWeights(const std::string& path, wetts_vits_config* cfg, int* sampling_rate,
                     std::vector<FlatTensor>* tensors, std::string* error) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { *error = "cannot open " + path; return false; }
  char magic[8];
  f.read(magic, 8);
  uint32_t version = 0, cfg_bytes = 0, n_tensors = 0;
  int32_t sr = 0;
  if (!f || std::memcmp(magic, kMagic, 8) != 0 || !ReadPod(f, &version) || version != 1 || !ReadPod(f, &cfg_bytes) ||
      cfg_bytes != sizeof(wetts_vits_config)) {
    *error = path + ": not a wetts_b200 flat weights file (or written for another ABI version)";
    return false;
  }
  f.read(reinterpret_cast<char*>(cfg), sizeof(*cfg));
  if (!ReadPod(f, &sr) || !ReadPod(f, &n_tensors)) { *error = path + ": truncated header"; return false; }
  *sampling_rate = sr;
  tensors->clear();
  // a malformed or hostile file must not drive allocation: every tensor has to fit in what is left of the file
  const std::streampos here = f.tellg();
  f.seekg(0, std::ios::end);
  const std::streamoff file_bytes = f.tellg();
  f.seekg(here);
  constexpr uint32_t kMaxTensors = 1u << 16;
  if (n_tensors > kMaxTensors) { *error = path + ": implausible tensor count"; return false; }
  for (uint32_t i = 0; i < n_tensors; ++i) {
    FlatTensor t;
    uint16_t name_len = 0;
    uint8_t ndim = 0;
    if (!ReadPod(f, &name_len)) { *error = path + ": truncated tensor table"; return false; }
    t.name.resize(name_len);
    f.read(&t.name[0], name_len);
    if (!ReadPod(f, &ndim) || ndim > 8) { *error = path + ": bad tensor rank"; return false; }
    size_t numel = 1;
    t.dims.resize(ndim);
    const size_t left = static_cast<size_t>(file_bytes - static_cast<std::streamoff>(f.tellg()));
    for (int d = 0; d < ndim; ++d) {
      if (!ReadPod(f, &t.dims[d]) || t.dims[d] <= 0) { *error = path + ": bad tensor shape"; return false; }
      const size_t dim = static_cast<size_t>(t.dims[d]);
      if (dim > left / sizeof(float) || numel > left / sizeof(float) / dim) {   // overflow-safe: numel*dim*4 <= left
        *error = path + ": tensor '" + t.name + "' is larger than the file";
        return false;
      }
      numel *= dim;
    }
    t.data.resize(numel);
    f.read(reinterpret_cast<char*>(t.data.data()), static_cast<std::streamsize>(numel * sizeof(float)));
    if (!f) { *error = path + ": truncated tensor '" + t.name + "'"; return false; }
    tensors->push_back(std::move(t));
  }
  return true;
}

void VitsModel::Check(int status, const char* what) const {
  if (status != 0) throw std::runtime_error(std::string(what) + ": " + wetts_last_error());
}

VitsModel::VitsModel(const std::string& weights_path, int chunk_size, int pad_size, int device, uint64_t seed)
    : rng_(seed), chunk_size_(chunk_size), pad_size_(pad_size) {
  std::vector<FlatTensor> tensors;
  std::string error;
  if (!LoadFlatWeights(weights_path, &cfg_, &sampling_rate_, &tensors, &error)) throw std::runtime_error(error);
  hidden_dim_ = cfg_.inter_channels;
  CudaCheck(cudaSetDevice(device), "cudaSetDevice");
  cudaStream_t s;
  CudaCheck(cudaStreamCreate(&s), "cudaStreamCreate");
  stream_ = s;
  Check(wetts_vits_create(&cfg_, device, &handle_), "wetts_vits_create");
  for (const FlatTensor& t : tensors)
    Check(wetts_vits_set_tensor(handle_, t.name.c_str(), t.data.data(), t.dims.data(), static_cast<int>(t.dims.size())),
          "wetts_vits_set_tensor");
  Check(wetts_vits_finalize(handle_), "wetts_vits_finalize");
  if (wetts_vits_upsample_factor(handle_) != kUpsampleRate)
    throw std::runtime_error("model upsample factor differs from kUpsampleRate (256)");
}

VitsModel::~VitsModel() {
  if (workspace_) cudaFree(workspace_);
  if (handle_) wetts_vits_destroy(handle_);
  if (stream_) cudaStreamDestroy(static_cast<cudaStream_t>(stream_));
}

void VitsModel::EnsureWorkspace(size_t bytes) {
  if (bytes <= workspace_bytes_) return;
  if (workspace_) cudaFree(workspace_);
  workspace_ = nullptr;
  workspace_bytes_ = 0;
  CudaCheck(cudaMalloc(&workspace_, bytes), "cudaMalloc(workspace)");
  workspace_bytes_ = bytes;
}

std::vector<float> VitsModel::ForwardEncoder(const std::vector<int64_t>& phonemes, int sid) {
  const int Tx = static_cast<int>(phonemes.size());
  if (Tx <= 0) throw std::invalid_argument("empty phoneme sequence");
  const int C = cfg_.inter_channels, H = cfg_.hidden_channels, gin = cfg_.gin_channels;
  const bool has_g = cfg_.n_speakers > 0 && gin > 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const float noise_scale = 0.667f, length_scale = 1.0f, noise_scale_w = 0.8f;   // vits_model.cc:51

  const int64_t len = Tx, spk = sid;
  DeviceBuffer d_ids(sizeof(int64_t) * Tx), d_len(sizeof(int64_t)), d_sid(sizeof(int64_t)), d_ylen(sizeof(int64_t));
  DeviceBuffer d_g(sizeof(float) * std::max(gin, 1)), d_h(sizeof(float) * H * Tx), d_m(sizeof(float) * C * Tx),
      d_logs(sizeof(float) * C * Tx), d_logw(sizeof(float) * Tx), d_wceil(sizeof(float) * Tx), d_cum(sizeof(int32_t) * Tx);
  CudaCheck(cudaMemcpyAsync(d_ids.p, phonemes.data(), sizeof(int64_t) * Tx, cudaMemcpyHostToDevice, s), "H2D ids");
  CudaCheck(cudaMemcpyAsync(d_len.p, &len, sizeof(int64_t), cudaMemcpyHostToDevice, s), "H2D length");
  CudaCheck(cudaMemcpyAsync(d_sid.p, &spk, sizeof(int64_t), cudaMemcpyHostToDevice, s), "H2D sid");
  if (has_g) Check(wetts_speaker_embedding(handle_, d_sid.as<int64_t>(), 1, d_g.as<float>(), s), "speaker embedding");
  const float* g = has_g ? d_g.as<float>() : nullptr;

  EnsureWorkspace(wetts_text_encoder_workspace_bytes(handle_, 1, Tx));
  Check(wetts_text_encoder_forward(handle_, d_ids.as<int64_t>(), d_len.as<int64_t>(), 1, Tx, d_h.as<float>(),
                                   d_m.as<float>(), d_logs.as<float>(), workspace_, workspace_bytes_, s),
        "text encoder");
  // stochastic duration predictor: explicit N(0,1) draws, as torch.randn in duration_predictors.py:257
  std::vector<float> noise_w;
  DeviceBuffer d_noise_w(sizeof(float) * 2 * Tx);
  if (cfg_.use_sdp) {
    std::normal_distribution<float> nd(0.f, 1.f);
    noise_w.resize(2 * static_cast<size_t>(Tx));
    for (float& v : noise_w) v = nd(rng_);
    CudaCheck(cudaMemcpyAsync(d_noise_w.p, noise_w.data(), sizeof(float) * noise_w.size(), cudaMemcpyHostToDevice, s),
              "H2D noise_w");
  }
  EnsureWorkspace(wetts_duration_workspace_bytes(handle_, 1, Tx));
  Check(wetts_duration_forward(handle_, d_h.as<float>(), d_len.as<int64_t>(), g,
                               cfg_.use_sdp ? d_noise_w.as<float>() : nullptr, noise_scale_w, 1, Tx, d_logw.as<float>(),
                               workspace_, workspace_bytes_, s),
        "duration predictor");
  Check(wetts_length_regulate(handle_, d_logw.as<float>(), d_len.as<int64_t>(), nullptr, length_scale, 1, Tx,
                              d_wceil.as<float>(), d_cum.as<int32_t>(), d_ylen.as<int64_t>(), s),
        "length regulation");
  int64_t Ty64 = 0;
  CudaCheck(cudaMemcpyAsync(&Ty64, d_ylen.p, sizeof(int64_t), cudaMemcpyDeviceToHost, s), "D2H y_length");
  CudaCheck(cudaStreamSynchronize(s), "stream sync");
  const int Ty = static_cast<int>(Ty64);

  // z_p = m_p + randn * exp(logs_p) * noise_scale (models.py:267), then the flow inversion
  std::vector<float> noise_z(static_cast<size_t>(C) * Ty);
  {
    std::normal_distribution<float> nd(0.f, 1.f);
    for (float& v : noise_z) v = nd(rng_);
  }
  DeviceBuffer d_noise_z(sizeof(float) * noise_z.size()), d_z(sizeof(float) * noise_z.size());
  CudaCheck(cudaMemcpyAsync(d_noise_z.p, noise_z.data(), sizeof(float) * noise_z.size(), cudaMemcpyHostToDevice, s),
            "H2D noise_z");
  Check(wetts_expand_prior(handle_, d_m.as<float>(), d_logs.as<float>(), d_cum.as<int32_t>(), d_len.as<int64_t>(),
                           d_ylen.as<int64_t>(), d_noise_z.as<float>(), static_cast<int64_t>(C) * Ty, Ty, noise_scale, 1,
                           Tx, Ty, nullptr, nullptr, d_z.as<float>(), nullptr, nullptr, s),
        "prior expansion");
  EnsureWorkspace(wetts_flow_workspace_bytes(handle_, 1, Ty));
  Check(wetts_flow_reverse(handle_, d_z.as<float>(), d_ylen.as<int64_t>(), g, 1, Ty, workspace_, workspace_bytes_, s),
        "flow inversion");
  std::vector<float> z_ct(static_cast<size_t>(C) * Ty);
  CudaCheck(cudaMemcpyAsync(z_ct.data(), d_z.p, sizeof(float) * z_ct.size(), cudaMemcpyDeviceToHost, s), "D2H z");
  CudaCheck(cudaStreamSynchronize(s), "stream sync");
  // channels-first [C][Ty] -> time-major [Ty][C] (the encoder graph returns z * y_mask transposed; B = 1: mask = 1)
  std::vector<float> z(static_cast<size_t>(Ty) * C);
  for (int c = 0; c < C; ++c)
    for (int t = 0; t < Ty; ++t) z[static_cast<size_t>(t) * C + c] = z_ct[static_cast<size_t>(c) * Ty + t];
  return z;
}

void VitsModel::ForwardDecoder(const std::vector<float>& z, int sid, std::vector<float>* audio) {
  const int C = cfg_.inter_channels;
  if (z.empty() || z.size() % static_cast<size_t>(C) != 0) throw std::invalid_argument("z must be [L][hidden_dim]");
  const int L = static_cast<int>(z.size() / C);
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const int64_t spk = sid;
  DeviceBuffer d_z(sizeof(float) * z.size()), d_sid(sizeof(int64_t)),
      d_audio(sizeof(float) * static_cast<size_t>(L) * kUpsampleRate);
  CudaCheck(cudaMemcpyAsync(d_z.p, z.data(), sizeof(float) * z.size(), cudaMemcpyHostToDevice, s), "H2D z");
  CudaCheck(cudaMemcpyAsync(d_sid.p, &spk, sizeof(int64_t), cudaMemcpyHostToDevice, s), "H2D sid");
  EnsureWorkspace(wetts_vits_decoder_workspace_bytes(handle_, 1, L));
  Check(wetts_vits_forward_decoder(handle_, d_z.as<float>(), cfg_.n_speakers > 0 ? d_sid.as<int64_t>() : nullptr, 1, L,
                                   d_audio.as<float>(), workspace_, workspace_bytes_, s),
        "decoder");
  audio->resize(static_cast<size_t>(L) * kUpsampleRate);
  CudaCheck(cudaMemcpyAsync(audio->data(), d_audio.p, sizeof(float) * audio->size(), cudaMemcpyDeviceToHost, s),
            "D2H audio");
  CudaCheck(cudaStreamSynchronize(s), "stream sync");
  for (float& v : *audio) v *= 32767.0f;   // vits_model.cc:84-86
}

void VitsModel::Forward(const std::vector<int64_t>& phonemes, int sid, std::vector<float>* audio) {
  ForwardDecoder(ForwardEncoder(phonemes, sid), sid, audio);
}

// inference_onnx.py:37-55 / vits_model.cc:96-111: ceil(L / chunk) chunks, each extended by `pad` frames on both sides
void VitsModel::SplitToChunks(const std::vector<float>& z) {
  if (chunk_size_ <= 0) throw std::invalid_argument("chunk_size must be positive for streaming");
  z_chunks_.clear();
  const int L = static_cast<int>(z.size() / hidden_dim_);
  const int num = (L + chunk_size_ - 1) / chunk_size_;
  for (int i = 0; i < num; ++i) {
    const int start = std::max(0, i * chunk_size_ - pad_size_);
    const int end = std::min((i + 1) * chunk_size_ + pad_size_, L);
    z_chunks_.emplace_back(z.begin() + static_cast<size_t>(start) * hidden_dim_,
                           z.begin() + static_cast<size_t>(end) * hidden_dim_);
  }
}

// inference_onnx.py:59-76 / vits_model.cc:114-125: drop the samples synthesised from the overlap frames
void VitsModel::Depadding(int chunk_id, int num_chunks, int chunk_size, int pad, int upsample, std::vector<float>* audio) {
  const size_t front = static_cast<size_t>(std::min(chunk_id * chunk_size, pad)) * upsample;
  const size_t body = static_cast<size_t>(chunk_size) * upsample;
  size_t lo = 0, hi = audio->size();
  if (chunk_id == 0) {
    hi = std::min(body, audio->size());
  } else if (chunk_id == num_chunks - 1) {
    lo = std::min(front, audio->size());
  } else {
    lo = std::min(front, audio->size());
    hi = std::min(front + body, audio->size());
  }
  audio->assign(audio->begin() + lo, audio->begin() + hi);
}

void VitsModel::SetInput(const std::vector<int64_t>& phonemes, int sid) {
  sid_ = sid;
  cur_ = 0;
  z_chunks_.clear();
  SplitToChunks(ForwardEncoder(phonemes, sid_));
}

bool VitsModel::StreamDecode(std::vector<float>* audio) {
  const int num_chunks = static_cast<int>(z_chunks_.size());
  if (cur_ < num_chunks) {
    ForwardDecoder(z_chunks_[cur_], sid_, audio);
    if (chunk_size_ > 0) Depadding(cur_, num_chunks, chunk_size_, pad_size_, kUpsampleRate, audio);
  }
  cur_++;   // at least one chunk inference per call (vits_model.cc:151)
  return cur_ >= num_chunks;
}

}  // namespace wetts
