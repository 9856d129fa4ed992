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
    W = len(plan)
    n_mel = spectrogram.shape[0]

    # ---- everything that does NOT depend on the previous window, for ALL windows at once -----------------------------
    # The window chain is serial only through the seed frames.  The audio encoder sees (spectrogram slice, speaker) and
    # the generators' text encoders see the tokens: both are evaluated here as batches over the clip's windows (eval-mode
    # BatchNorm is per-sample, so a window's features do not depend on its batch mates) instead of once per window at
    # batch 1 inside the loop -- the loop body keeps only the recurrent decoders.
    starts = [p[2] for p in plan]
    if starts[-1] + spec_len > spectrogram.shape[1]:   # the reference would fail on a short slice too; keep the error explicit
        raise RuntimeError(f"spectrogram window {W - 1} has {spectrogram.shape[1] - starts[-1]} frames, expected {spec_len}")
    win_idx = (torch.tensor(starts, device=dev).view(W, 1) + torch.arange(spec_len, device=dev).view(1, spec_len)).reshape(-1)
    spec_all = spectrogram.index_select(1, win_idx).reshape(n_mel, W, spec_len).permute(1, 0, 2).contiguous()   # [W,128,70]
    text_all = torch.from_numpy(np.stack([place_words(words, p[0], p[1], n_frames, lang_model) for p in plan])).to(dev)
    a0 = math.floor(plan[-1][0] / clip_length * len(audio))
    if len(audio) - a0 < audio_sample_length:
        end_padding_duration = audio_sample_length - max(0, len(audio) - a0)
    vid_all = vid_t.expand(W).contiguous()
    CH = 128                                            # windows per encoder batch (bounds the activation memory)
    blends_all = [[] for _ in range(L)]
    text_feat_all = [[] for _ in range(L)]
    for c0 in range(0, W, CH):
        sl = slice(c0, min(W, c0 + CH))
        _, _, _, _, bl = audio_encoder(spec_all[sl], vid_all[sl])
        for k in range(L):
            blends_all[k].append(bl[k])
            text_feat_all[k].append(gens[k].text_encoder(text_all[sl]))
    # [W, 2L, 34, 32]: per window the L blended audio features then the L text features (one copy per window below)
    feats_all = torch.stack([torch.cat(x) for x in blends_all] + [torch.cat(x) for x in text_feat_all], dim=1).contiguous()

    # static buffers of the window body (also what a captured graph reads and writes)
    s_feats = torch.empty((2 * L, 1, n_frames, 32), device=dev, dtype=torch.float32)
    s_blend = [s_feats[k] for k in range(L)]
    s_tfeat = [s_feats[L + k] for k in range(L)]
    s_text = torch.zeros((1, n_frames), device=dev, dtype=torch.int64)   # (unused by the generators once _text_feat is given)
    s_out = torch.zeros((1, n_frames, pose_dim), device=dev, dtype=torch.float32)
    out_all = torch.empty((W, n_frames, pose_dim), device=dev, dtype=torch.float32)

    def body(with_seed: bool):
        if with_seed:  # seed frames: the previous window's last n_pre outputs, per level (:117-125)
            seed = s_out[:, -n_pre:, :].contiguous()
            for k in range(L):
                targets[k][:, 0:n_pre, :] = ops.gather_cols(seed, tabs[k][0])
        outs, _ = cascade.run_cascade(variant, gens, targets, s_text, s_blend, vid_t, n_pre, text_feats=s_tfeat)
        s_out.copy_(outs[-1])

    use_graph = _GRAPH and not rng.overridden() and W >= 4 and not torch.cuda.is_current_stream_capturing()
    graph = None
    for i in range(W):
        s_feats.copy_(feats_all[i].unsqueeze(1))
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
        out_all[i].copy_(s_out[0])

    # linear cross-fade over the overlapping n_pre frames (:195-203), all windows at once on the device: window i keeps
    # its frames [0, n_frames - n_pre) (the last window all of them) and its first n_pre frames are blended with the
    # previous window's last n_pre
    ops._call("ha2g_crossfade", ops._p(out_all), W, n_frames, pose_dim, n_pre, ops._st())
    chunks = [out_all[:-1, :n_frames - n_pre].reshape(-1, pose_dim), out_all[-1]]

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
