"""Window-level predictor: the fused replacement of the reference's per-batch loop body
(helen/modules/python/models/predict_gpu.py:97-159) behind the C ABI."""
import ctypes
from collections import OrderedDict

import numpy as np
import torch

from . import _native
from .options import ImageSizeOptions, TrainOptions

STATE_DICT_KEYS = tuple(
    f"{layer}.{name}{rev}"
    for layer in ("gru_encoder", "gru_decoder")
    for rev in ("", "_reverse")
    for name in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")
) + ("dense1_base.weight", "dense1_base.bias", "dense2_rle.weight", "dense2_rle.bias")


def normalise_state_dict(state_dict):
    """Strip the DataParallel/DDP ``module.`` prefix (ModelHander.py:69-74) and return
    contiguous fp32 CPU tensors keyed like TransducerGRU.state_dict()."""
    out = OrderedDict()
    for key, value in state_dict.items():
        name = key[7:] if key[0:7] == "module." else key
        if not torch.is_tensor(value):
            value = torch.as_tensor(np.asarray(value))
        out[name] = value.detach().to(device="cpu", dtype=torch.float32).contiguous()
    missing = [k for k in STATE_DICT_KEYS if k not in out]
    if missing:
        raise KeyError(f"state_dict is missing parameters: {missing}")
    return out


def _fptr(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


class WindowPredictor(object):
    """Owns one native handle (packed weights on one GPU) plus a reusable device workspace."""

    def __init__(self, state_dict, device=0, engine="default"):
        lib = _native.load()
        sd = normalise_state_dict(state_dict)
        hidden = sd["gru_encoder.weight_hh_l0"].shape[1]
        features = sd["gru_encoder.weight_ih_l0"].shape[1]
        if sd["gru_encoder.weight_hh_l0"].shape[0] != 3 * hidden:
            raise ValueError("weight_hh_l0 is not [3H, H]")
        w = _native.hb_weights()
        for layer, dst in (("gru_encoder", w.encoder), ("gru_decoder", w.decoder)):
            for d, rev in enumerate(("", "_reverse")):
                dst.weight_ih[d] = _fptr(sd[f"{layer}.weight_ih_l0{rev}"])
                dst.weight_hh[d] = _fptr(sd[f"{layer}.weight_hh_l0{rev}"])
                dst.bias_ih[d] = _fptr(sd[f"{layer}.bias_ih_l0{rev}"])
                dst.bias_hh[d] = _fptr(sd[f"{layer}.bias_hh_l0{rev}"])
        w.base_weight, w.base_bias = _fptr(sd["dense1_base.weight"]), _fptr(sd["dense1_base.bias"])
        w.rle_weight, w.rle_bias = _fptr(sd["dense2_rle.weight"]), _fptr(sd["dense2_rle.bias"])
        if isinstance(device, torch.device):
            device = device.index or 0
        self._lib = lib
        self._handle = ctypes.c_void_p()
        _native.check(lib.hb_create(ctypes.byref(w), features, hidden, sd["dense1_base.weight"].shape[0],
                                    sd["dense2_rle.weight"].shape[0], int(device), ctypes.byref(self._handle)))
        self.device = torch.device("cuda", int(device))
        self.image_features = features
        self.hidden_size = hidden
        self._workspace = None
        if engine != "default":
            self.set_engine(engine)

    # -- lifecycle ------------------------------------------------------------------------
    def close(self):
        handle, self._handle = getattr(self, "_handle", None), None     # forgotten before it is freed (close() also runs at shutdown)
        if handle is not None and handle.value:
            self._lib.hb_destroy(handle)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration --------------------------------------------------------------------
    def set_engine(self, engine):
        _native.check(self._lib.hb_set_engine(self._handle, _native.ENGINES[engine]))

    @property
    def engine(self):
        code = _native.check(self._lib.hb_get_engine(self._handle))
        return {_native.ENGINE_FP32: "fp32", _native.ENGINE_TENSOR: "tensor"}[code]

    @property
    def launch_count(self):
        return int(self._lib.hb_launch_count(self._handle))

    def enable_kernel_timing(self, enable=True):
        """True/1: bracket the whole launch sequence of each predict call; 2: also each launch of the dominant kernel."""
        _native.check(self._lib.hb_enable_kernel_timing(self._handle, int(enable)))

    def last_launch_plan(self):
        """How the last predict call was laid out on the chip (hb_last_launch_plan)."""
        plan = _native.hb_launch_plan()
        _native.check(self._lib.hb_last_launch_plan(self._handle, ctypes.byref(plan)))
        return {name: int(getattr(plan, name)) for name, _ in plan._fields_}

    def dominant_kernel_time_ms(self, reset=True):
        total, n = ctypes.c_double(), ctypes.c_int64()
        _native.check(self._lib.hb_dominant_kernel_time_ms(self._handle, ctypes.byref(total), ctypes.byref(n), int(reset)))
        return total.value, n.value

    def kernel_time_ms(self, reset=True):
        total, n = ctypes.c_double(), ctypes.c_int64()
        _native.check(self._lib.hb_kernel_time_ms(self._handle, ctypes.byref(total), ctypes.byref(n), int(reset)))
        return total.value, n.value

    def workspace_bytes(self, batch, seq_len, window):
        out = ctypes.c_size_t()
        _native.check(self._lib.hb_workspace_bytes(self._handle, batch, seq_len, window, ctypes.byref(out)))
        return out.value

    def _get_workspace(self, batch, seq_len, window):
        need = self.workspace_bytes(batch, seq_len, window)
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    # -- compute --------------------------------------------------------------------------
    def predict(self, images, return_probs=False, window=TrainOptions.TRAIN_WINDOW, jump=TrainOptions.WINDOW_JUMP):
        """images: uint8 CUDA tensor [B, T, F].  Returns (base_labels u8 [B,T], rle_labels u8 [B,T])
        and, if return_probs, the accumulated softmax sums (base [B,T,5], rle [B,T,11])."""
        if not torch.is_tensor(images) or not images.is_cuda:
            raise ValueError("predict() takes a CUDA tensor; use predict_host() for host arrays (no CPU path exists)")
        if images.dtype != torch.uint8 or images.dim() != 3 or images.shape[2] != self.image_features:
            raise ValueError(f"images must be uint8 [B, T, {self.image_features}], got {images.dtype} {tuple(images.shape)}")
        if images.device != self.device:
            raise ValueError(f"images on {images.device}, predictor on {self.device}")
        images = images.contiguous()
        batch, seq_len = images.shape[0], images.shape[1]
        # the library writes every element of its outputs (columns no chunk covers get label 0 / probability 0):
        # no fill kernels here
        alloc = torch.empty if (batch and seq_len) else torch.zeros
        base = alloc((batch, seq_len), dtype=torch.uint8, device=self.device)
        rle = alloc((batch, seq_len), dtype=torch.uint8, device=self.device)
        pb = pr = None
        if return_probs:
            pb = alloc((batch, seq_len, ImageSizeOptions.TOTAL_BASE_LABELS), dtype=torch.float32, device=self.device)
            pr = alloc((batch, seq_len, ImageSizeOptions.TOTAL_RLE_LABELS), dtype=torch.float32, device=self.device)
        if batch and seq_len:
            ws = self._get_workspace(batch, seq_len, window)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _native.check(self._lib.hb_predict_windows(
                self._handle, images.data_ptr(), batch, seq_len, window, jump, base.data_ptr(), rle.data_ptr(),
                pb.data_ptr() if return_probs else None, pr.data_ptr() if return_probs else None,
                ws.data_ptr(), ws.numel(), stream))
        return (base, rle, pb, pr) if return_probs else (base, rle)

    def predict_host(self, images, return_probs=False, window=TrainOptions.TRAIN_WINDOW, jump=TrainOptions.WINDOW_JUMP):
        """images: uint8 numpy array / CPU tensor [B, T, F].  Host in, host out (numpy)."""
        if torch.is_tensor(images):
            images = images.cpu().numpy()
        images = np.ascontiguousarray(images)
        if images.dtype != np.uint8 or images.ndim != 3 or images.shape[2] != self.image_features:
            raise ValueError(f"images must be uint8 [B, T, {self.image_features}], got {images.dtype} {images.shape}")
        batch, seq_len = images.shape[0], images.shape[1]
        base = np.zeros((batch, seq_len), np.uint8)
        rle = np.zeros((batch, seq_len), np.uint8)
        pb = np.zeros((batch, seq_len, ImageSizeOptions.TOTAL_BASE_LABELS), np.float32) if return_probs else None
        pr = np.zeros((batch, seq_len, ImageSizeOptions.TOTAL_RLE_LABELS), np.float32) if return_probs else None
        _native.check(self._lib.hb_predict_windows_host(
            self._handle, images.ctypes.data, batch, seq_len, window, jump, base.ctypes.data, rle.ctypes.data,
            pb.ctypes.data if return_probs else None, pr.ctypes.data if return_probs else None))
        return (base, rle, pb, pr) if return_probs else (base, rle)

    def forward_chunk(self, x, hidden):
        """TransducerGRU.forward for one chunk: x f32 CUDA [B, W, F], hidden f32 CUDA [B, 2, H]."""
        if not (torch.is_tensor(x) and x.is_cuda and torch.is_tensor(hidden) and hidden.is_cuda):
            raise ValueError("forward_chunk() takes CUDA tensors (no CPU path exists)")
        if x.dim() != 3 or x.shape[2] != self.image_features:
            raise ValueError(f"x must be [B, W, {self.image_features}], got {tuple(x.shape)}")
        batch, width = x.shape[0], x.shape[1]
        if tuple(hidden.shape) != (batch, 2, self.hidden_size):
            raise ValueError(f"hidden must be [{batch}, 2, {self.hidden_size}], got {tuple(hidden.shape)}")
        x = x.to(torch.float32).contiguous()
        hidden = hidden.to(torch.float32).contiguous()
        base = torch.empty((batch, width, ImageSizeOptions.TOTAL_BASE_LABELS), dtype=torch.float32, device=self.device)
        rle = torch.empty((batch, width, ImageSizeOptions.TOTAL_RLE_LABELS), dtype=torch.float32, device=self.device)
        h_out = torch.empty_like(hidden)
        if batch and width:
            ws = self._get_workspace(batch, width, width)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _native.check(self._lib.hb_forward_chunk(
                self._handle, x.data_ptr(), hidden.data_ptr(), batch, width, base.data_ptr(), rle.data_ptr(),
                h_out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return base, rle, h_out
