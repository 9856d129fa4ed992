"""Sliding-window autoregressive inference on the GPU: drop-in for ``generate_gestures_hierarchy``
(scripts/synthesize_expressive_hierarchy.py:36-259; the TED-Gesture twin is synthesize_hierarchy.py:36-215).

Same signature and return value (numpy [num_frames, pose_dim]); the loop body -- log-mel of the whole clip,
per-window spectrogram slice, word-to-frame placement, seed frames from the previous window's last
``n_pre_poses`` outputs, audio encoder + L-level cascade at batch 1, linear cross-fade over the overlap -- follows
the reference line by line, including its index quirk (the spectrogram window start is computed from
``spectrogram.shape[0]`` = 128 mel bins where the time length was meant, :84).  What changes is where it runs: the
mel spectrogram is csrc/mel.cu, the modules are the CUDA modules of ha2g_b200.model, the seed frames stay on the
device between windows and the cross-fade runs on the device; only the final stacked motion is copied to the host.
The window body (seed frames -> audio encoder -> L-level cascade, ~600 launches at batch 1) is captured into one CUDA
graph at the second window and replayed for the rest of the clip (``HA2G_CUDA_GRAPH=0``, or injected randomness, keeps
every window eager).
"""
from __future__ import annotations

import math
import random
from typing import List, Sequence

import numpy as np
import torch

import os

from . import cascade, mel, ops, rng

_GRAPH = os.environ.get("HA2G_CUDA_GRAPH", "1") != "0"


def words_in_time_range(word_list: Sequence, start_time: float, end_time: float) -> List:
    """DataPreprocessor.get_words_in_time_range (scripts/data_loader/data_preprocessor_expressive.py:174-188)."""
    words = []
    for word in word_list:
        _, word_s, word_e = word[0], word[1], word[2]
        if word_s >= end_time:
            break
        if word_e <= start_time:
            continue
        words.append(word)
    return words


def window_plan(n_samples: int, audio_sr: int, n_poses: int, n_pre_poses: int, fps: float, n_mel_rows: int = 128):
    """Index arithmetic of the window loop (bit-exact contract): per window (start_time, end_time, spec_start)."""
    clip_length = n_samples / audio_sr
    unit_time = n_poses / fps
    stride_time = (n_poses - n_pre_poses) / fps
    if clip_length < unit_time:
        num_subdivision = 1
    else:
        num_subdivision = math.ceil((clip_length - unit_time) / stride_time) + 1
    plan = []
    for i in range(num_subdivision):
        start_time = i * stride_time
        plan.append((start_time, start_time + unit_time, math.floor(start_time / clip_length * n_mel_rows)))
    return plan


def place_words(words, start_time: float, end_time: float, n_frames: int, lang_model) -> np.ndarray:
    """extended_word_indices of one window (synthesize_expressive_hierarchy.py:101-111)."""
    ext = np.zeros(n_frames)
    frame_duration = (end_time - start_time) / n_frames
    for word in words_in_time_range(words, start_time, end_time):
        idx = max(0, int(np.floor((word[1] - start_time) / frame_duration)))
        ext[idx] = lang_model.get_word_index(word[0])
    return ext.astype(np.int64)


@torch.no_grad()
def generate_gestures_hierarchy(args, *rest, audio_sr=16000, vid=None, fade_out=False):
    """generate_gestures_hierarchy(args, g1..gL, audio_encoder, lang_model, audio, words, target_1..target_L, ...)

    L = 6 (TED-Expressive) or 3 (TED-Gesture) is inferred from the argument count, so the one function serves both
    reference scripts.  ``audio`` is a 1-D float array at 16 kHz, ``words`` a list of (word, start_s, end_s)."""
    L = (len(rest) - 4) // 2
    if L not in (3, 6) or len(rest) != 2 * L + 4:
        raise TypeError("expected (args, g1..gL, audio_encoder, lang_model, audio, words, target_1..target_L)")
    variant = "expressive" if L == 6 else "gesture"
    gens = list(rest[:L])
    audio_encoder, lang_model, audio, words = rest[L:L + 4]
    targets = [t.clone() for t in rest[L + 4:]]
    dev = next(gens[0].parameters()).device
    n_frames, n_pre = args.n_poses, args.n_pre_poses
    fps = args.motion_resampling_framerate
    pose_dim = len(args.mean_dir_vec)

    if torch.is_tensor(audio):   # already a tensor (pinned host or device-resident clip): no numpy round trip
        audio_t = audio.to(device=dev, dtype=torch.float32, non_blocking=True)
    else:
        audio_t = torch.as_tensor(np.asarray(audio), dtype=torch.float32).to(dev)
    spectrogram = mel.extract_melspectrogram(audio_t)  # [128, frames] on the device
    plan = window_plan(len(audio), audio_sr, n_frames, n_pre, fps, spectrogram.shape[0])
    spec_len = mel.calc_spectrogram_length_from_motion_length(n_frames, fps)
    audio_sample_length = int(n_frames / fps * audio_sr)
    clip_length = len(audio) / audio_sr
    end_padding_duration = 0

    if args.z_type == "speaker":
        if not vid:
            vid = random.randrange(gens[0].z_obj.n_words)
        vid_t = torch.tensor([vid], dtype=torch.int64, device=dev)
    else:
        raise NotImplementedError("z_type must be 'speaker' on the hierarchy path")

    targets = [t.to(dev).float() for t in targets]
    tabs = cascade.device_tables(variant, dev)
    # static buffers of the window body (also what a captured graph reads and writes)
    s_spec = torch.empty((1, spectrogram.shape[0], spec_len), device=dev, dtype=torch.float32)
    s_text = torch.empty((1, n_frames), device=dev, dtype=torch.int64)
    s_out = torch.zeros((1, n_frames, pose_dim), device=dev, dtype=torch.float32)

    def body(with_seed: bool):
        if with_seed:  # seed frames: the previous window's last n_pre outputs, per level (:117-125)
            seed = s_out[:, -n_pre:, :].contiguous()
            for k in range(L):
                targets[k][:, 0:n_pre, :] = ops.gather_cols(seed, tabs[k][0])
        _, _, _, _, linear_blend_feat = audio_encoder(s_spec, vid_t)
        outs, _ = cascade.run_cascade(variant, gens, targets, s_text, linear_blend_feat, vid_t, n_pre)
        s_out.copy_(outs[-1])

    use_graph = _GRAPH and not rng.overridden() and len(plan) >= 4 and not torch.cuda.is_current_stream_capturing()
    graph = None
    chunks: List[torch.Tensor] = []
    for i, (start_time, end_time, spec_start) in enumerate(plan):
        sl = spectrogram[:, spec_start:spec_start + spec_len]
        if sl.shape[1] == spec_len:
            s_spec[0].copy_(sl)
        else:   # the reference would fail on a short slice too; keep the error explicit
            raise RuntimeError(f"spectrogram window {i} has {sl.shape[1]} frames, expected {spec_len}")
        a0 = math.floor(start_time / clip_length * len(audio))
        if len(audio) - a0 < audio_sample_length and i == len(plan) - 1:
            end_padding_duration = audio_sample_length - max(0, len(audio) - a0)
        s_text.copy_(torch.from_numpy(place_words(words, start_time, end_time, n_frames, lang_model)).unsqueeze(0))
        if use_graph and i == 2:   # windows 0 and 1 ran eagerly (both variants of the body are warm): capture the seeded body
            ops._ensure_workspace()
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body(True)
        if graph is not None:
            graph.replay()
        else:
            body(i > 0)
        out_seq = s_out[0].clone()
        if chunks:  # linear cross-fade over the overlapping n_pre frames (:195-203)
            last = chunks[-1][-n_pre:]
            chunks[-1] = chunks[-1][:-n_pre]
            n = last.shape[0]
            j = torch.arange(n, device=dev, dtype=torch.float32).unsqueeze(1)
            out_seq[:n] = last * (n - j) / (n + 1) + out_seq[:n] * (j + 1) / (n + 1)
        chunks.append(out_seq)

    result = torch.cat(chunks, dim=0).cpu().numpy()

    if fade_out:  # host-side post-processing exactly as the reference (:238-257)
        n_smooth = n_pre
        start_frame = len(result) - int(end_padding_duration / audio_sr * fps)
        end_frame = start_frame + n_smooth * 2
        if len(result) < end_frame:
            result = np.pad(result, [(0, end_frame - len(result)), (0, 0)], mode="constant")
        result[end_frame - n_smooth:] = np.zeros((pose_dim,))
        y = result[start_frame:end_frame]
        x = np.array(range(0, y.shape[0]))
        w = np.ones(len(y))
        w[0] = 5
        w[-1] = 5
        coeffs = np.polyfit(x, y, 2, w=w)
        fit = [np.poly1d(coeffs[:, k]) for k in range(0, y.shape[1])]
        result[start_frame:end_frame] = np.transpose(np.asarray([f(x) for f in fit]))
    return result
