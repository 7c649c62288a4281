// Command-line driver of the C++ VitsModel shim (the part of runtime/bin/tts_main.cc:80-101 behind the text
// front-end): phoneme ids -> wav, non-streaming or chunked streaming.
//
//   vits_main --weights model.wb2 (--phonemes "12 7 33 ..." | --decode_z z.f32) [--sid 0] [--stream]
//             [--chunk 40] [--pad 10] [--seed 1234] [--wav out.wav] [--f32 out.f32]
//
// --decode_z reads a time-major z [L][hidden] float32 file and runs only the decoder (ForwardDecoder / the
// streaming decode), which makes the output comparable bit for bit with the Python session adapters.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "vits_model.h"

namespace {

void WriteWav16(const std::string& path, const std::vector<float>& audio, int sample_rate) {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot write " + path);
  const uint32_t n = static_cast<uint32_t>(audio.size()), data_bytes = n * 2, riff = 36 + data_bytes, sr = sample_rate;
  const uint16_t fmt = 1, ch = 1, bits = 16, align = 2;
  const uint32_t fmt_len = 16, byte_rate = sr * 2;
  f.write("RIFF", 4); f.write(reinterpret_cast<const char*>(&riff), 4); f.write("WAVEfmt ", 8);
  f.write(reinterpret_cast<const char*>(&fmt_len), 4); f.write(reinterpret_cast<const char*>(&fmt), 2);
  f.write(reinterpret_cast<const char*>(&ch), 2); f.write(reinterpret_cast<const char*>(&sr), 4);
  f.write(reinterpret_cast<const char*>(&byte_rate), 4); f.write(reinterpret_cast<const char*>(&align), 2);
  f.write(reinterpret_cast<const char*>(&bits), 2); f.write("data", 4);
  f.write(reinterpret_cast<const char*>(&data_bytes), 4);
  for (float v : audio) {   // samples are already scaled by 32767 (vits_model.cc:84-86); clip as frontend/wav.h does
    const float c = v > 32767.f ? 32767.f : (v < -32768.f ? -32768.f : v);
    const int16_t q = static_cast<int16_t>(c);
    f.write(reinterpret_cast<const char*>(&q), 2);
  }
}

int Usage(const char* argv0) {
  std::fprintf(stderr,
               "usage: %s --weights FILE (--phonemes \"ids\" | --decode_z FILE) [--sid N] [--stream] [--chunk N] [--pad N]\n"
               "          [--seed N] [--wav FILE] [--f32 FILE]\n",
               argv0);
  return 64;
}

}  // namespace

int main(int argc, char** argv) {
  std::string weights, phonemes, z_path, wav, f32;
  int sid = 0, chunk = 40, pad = 10;
  unsigned long long seed = 1234;
  bool stream = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : ""; };
    if (a == "--weights") weights = next();
    else if (a == "--phonemes") phonemes = next();
    else if (a == "--decode_z") z_path = next();
    else if (a == "--sid") sid = std::atoi(next());
    else if (a == "--chunk") chunk = std::atoi(next());
    else if (a == "--pad") pad = std::atoi(next());
    else if (a == "--seed") seed = std::strtoull(next(), nullptr, 10);
    else if (a == "--wav") wav = next();
    else if (a == "--f32") f32 = next();
    else if (a == "--stream") stream = true;
    else return Usage(argv[0]);
  }
  if (weights.empty() || (phonemes.empty() == z_path.empty())) return Usage(argv[0]);
  try {
    wetts::VitsModel model(weights, chunk, pad, /*device=*/0, seed);
    std::vector<float> z, audio;
    std::vector<int64_t> ids;
    if (!z_path.empty()) {
      std::ifstream f(z_path, std::ios::binary | std::ios::ate);
      if (!f) throw std::runtime_error("cannot open " + z_path);
      const std::streamsize bytes = f.tellg();
      f.seekg(0);
      z.resize(static_cast<size_t>(bytes) / sizeof(float));
      f.read(reinterpret_cast<char*>(z.data()), bytes);
    } else {
      std::istringstream is(phonemes);
      long long v;
      while (is >> v) ids.push_back(v);
      z = model.ForwardEncoder(ids, sid);
    }
    if (!stream) {
      model.ForwardDecoder(z, sid, &audio);
    } else {   // SetInput + StreamDecode loop of tts_main.cc, on the z computed above
      model.SplitToChunks(z);
      const int n = model.num_chunks();
      for (int c = 0; c < n; ++c) {
        std::vector<float> piece;
        // same calls StreamDecode makes (vits_model.cc:139-150), driven explicitly so that --decode_z can stream too
        std::vector<float> chunk_z(z.begin() + static_cast<size_t>(std::max(0, c * chunk - pad)) * model.hidden_dim(),
                                   z.begin() + static_cast<size_t>(std::min((c + 1) * chunk + pad,
                                                                           static_cast<int>(z.size() / model.hidden_dim()))) *
                                                   model.hidden_dim());
        model.ForwardDecoder(chunk_z, sid, &piece);
        model.Depadding(c, n, chunk, pad, wetts::kUpsampleRate, &piece);
        audio.insert(audio.end(), piece.begin(), piece.end());
      }
    }
    if (!f32.empty()) {
      std::ofstream f(f32, std::ios::binary);
      f.write(reinterpret_cast<const char*>(audio.data()), static_cast<std::streamsize>(audio.size() * sizeof(float)));
    }
    if (!wav.empty()) WriteWav16(wav, audio, model.sampling_rate());
    std::printf("frames %zu samples %zu sampling_rate %d\n", z.size() / model.hidden_dim(), audio.size(), model.sampling_rate());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "vits_main: %s\n", e.what());
    return 1;
  }
  return 0;
}
