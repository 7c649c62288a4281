"""L2 adapters: the ONNX-session contract, streaming chunk helpers and
the int16 output stage, on top of the CUDA engine.

* `InferenceSession` has `onnxruntime.InferenceSession.run(None, feeds)` semantics for the three
  graphs the reference exports (wetts/vits/export_onnx.py:93-189): the full model
  {"input","input_lengths","scales","sid"} -> [output f32[B,1,L]], the streaming encoder (same
  feeds -> [z f32[B,L,192]]) and decoder ({"z","sid"} -> [output]).  It drops into
  wetts/cli/model.py:27-28,48-59 and wetts/vits/inference_onnx.py:129-158 unchanged.
* `split_to_chunks` / `depadding` / `StreamingVits` restate the chunked vocoding of
  inference_onnx.py:37-76 and runtime/core/model/vits_model.cc:96-153 (VitsModel::SetInput /
  StreamDecode): block `chunk_size`, overlap `pad_size`, kUpsampleRate = 256.
* `to_int16` is the callers' output stage, done on the device: plain x32767 (cli/model.py:60,
  vits_model.cc:84-86), per-utterance peak-normalise x0.6 (inference.py:101-105) or the
  batch-global variant of runtime/gpu_triton/model_repo/tts/1/model.py:150-151.
"""
import math

import numpy as np
import torch


class _IO:
    def __init__(self, name):
        self.name = name


class InferenceSession:
    """`run(output_names, feeds)` -> list of numpy arrays, like onnxruntime."""

    GRAPHS = {
        "full": (["input", "input_lengths", "scales", "sid"], ["output"]),
        "encoder": (["input", "input_lengths", "scales", "sid"], ["z"]),
        "decoder": (["z", "sid"], ["output"]),
    }

    def __init__(self, net, graph="full"):
        if graph not in self.GRAPHS:
            raise ValueError(f"graph must be one of {sorted(self.GRAPHS)}")
        self.net, self.graph = net, graph

    def get_inputs(self):
        return [_IO(n) for n in self.GRAPHS[self.graph][0]]

    def get_outputs(self):
        return [_IO(n) for n in self.GRAPHS[self.graph][1]]

    def run(self, output_names, feeds):
        need = self.GRAPHS[self.graph][0]
        missing = [n for n in need if n not in feeds]
        if missing:
            raise ValueError(f"missing feeds {missing} for the {self.graph} graph")  # ORT raises here too
        t = {k: torch.as_tensor(np.asarray(v)) for k, v in feeds.items()}
        if self.graph == "decoder":
            out = self.net.export_decoder_forward(t["z"].float(), t["sid"].long())
        else:
            scales = t["scales"].float().reshape(-1, 3)
            fn = self.net.export_forward if self.graph == "full" else self.net.export_encoder_forward
            out = fn(t["input"].long(), t["input_lengths"].long(), scales, t["sid"].long())
        return [out.detach().cpu().numpy()]


def split_to_chunks(z, chunk_size, pad_size):
    """z [B, L, C] (time-major, the encoder graph's output) -> list of overlapping chunks
    (inference_onnx.py:37-55, vits_model.cc:96-111).  chunk_size == -1: one chunk."""
    if chunk_size == -1:
        return [z]
    L = z.shape[1]
    n = math.ceil(L / chunk_size)
    return [z[:, max(0, i * chunk_size - pad_size): min((i + 1) * chunk_size + pad_size, L), :] for i in range(n)]


def depadding(audio, chunk_num, chunk_id, block, pad, upsample=256):
    """Drop the samples synthesised from the overlap (inference_onnx.py:59-76, vits_model.cc:114-125).
    audio [B, T]."""
    assert audio.dim() == 2 if torch.is_tensor(audio) else audio.ndim == 2
    front = min(chunk_id * block, pad)
    if chunk_id == 0:
        return audio[:, : block * upsample]
    if chunk_id == chunk_num - 1:
        return audio[:, front * upsample:]
    return audio[:, front * upsample: (front + block) * upsample]


class StreamingVits:
    """VitsModel::SetInput / StreamDecode (vits_model.cc:127-153): encode once, vocode chunk by chunk."""

    def __init__(self, net, chunk_size=40, pad_size=10, scales=(0.667, 1.0, 0.8)):
        self.net, self.chunk_size, self.pad_size = net, chunk_size, pad_size
        self.scales = torch.tensor([list(scales)], dtype=torch.float32)
        self.chunks, self.cur, self.sid = [], 0, None

    def set_input(self, phonemes, sid):
        x = torch.as_tensor(phonemes, dtype=torch.long).reshape(1, -1)
        lens = torch.tensor([x.shape[1]])
        self.sid = torch.tensor([int(sid)])
        z = self.net.export_encoder_forward(x, lens, self.scales, self.sid)
        self.chunks = split_to_chunks(z, self.chunk_size, self.pad_size)
        self.cur = 0

    def stream_decode(self):
        """Returns (audio_chunk f32[T] on the device scaled to int16 range like ForwardDecoder, done)."""
        n = len(self.chunks)
        audio = None
        if self.cur < n:
            a = self.net.export_decoder_forward(self.chunks[self.cur], self.sid)[:, 0]
            if self.chunk_size > 0:
                a = depadding(a, n, self.cur, self.chunk_size, self.pad_size, self.net._engine.upsample)
            audio = a[0] * 32767.0
        self.cur += 1
        return audio, self.cur >= n


def to_int16(audio, mode="scale", lengths=None):
    """audio f32[B,1,L] or [B,L] in [-1,1] -> int16 (same shape), on the audio's device (one CUDA kernel pair behind
    `wetts_audio_to_int16`, no CPU path).  mode "scale": x32767; "peak": per utterance
    32767/max(0.01, max|a|)*0.6 then clip; "peak_batch": one gain for the whole batch.  `lengths` (valid samples
    per row) restricts the peak search of "peak" to each utterance's own samples."""
    from . import _lib
    from ._lib import WettsError, check
    modes = {"scale": 0, "peak": 1, "peak_batch": 2}
    if mode not in modes:
        raise ValueError(mode)
    if not (torch.is_tensor(audio) and audio.is_cuda):
        raise WettsError("to_int16 runs on CUDA tensors only (wetts_b200 has no CPU fallback)")
    a = audio.float().contiguous()
    B = a.shape[0]
    L = a.numel() // B
    out = torch.empty(a.shape, dtype=torch.int16, device=a.device)
    scratch = torch.empty(max(B, 1), dtype=torch.float32, device=a.device)
    ln = None if lengths is None else lengths.to(device=a.device, dtype=torch.int64).contiguous()
    with torch.cuda.device(a.device):
        st = torch.cuda.current_stream(a.device).cuda_stream
        check(_lib.load().wetts_audio_to_int16(a.data_ptr(), None if ln is None else ln.data_ptr(), B, L, modes[mode],
                                               scratch.data_ptr(), out.data_ptr(), st))
    return out
