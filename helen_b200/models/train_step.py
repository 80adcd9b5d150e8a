"""One optimisation step over one chunk: the body of the reference's training loop
(helen/modules/python/models/train.py:174-206) behind the C ABI.

    trainer = ChunkTrainer(model, device=0)                  # model: helen_b200 TransducerGRU on CUDA
    for i in range(0, SEQ_LENGTH, WINDOW_JUMP):               # train.py:174
        model_optimizer.zero_grad()                           # :176
        ...
        loss, loss_base, loss_rle, hidden = trainer.step(image_chunk, hidden, label_base_chunk, label_rle_chunk)
        model_optimizer.step()                                # :202  (torch.optim.Adam, untouched)

``step`` replaces :189 (forward), :192-198 (CrossEntropyLoss + CrossEntropyLoss(weight=CLASS_WEIGHTS)) and :201
(loss.backward()): it fills ``p.grad`` of every parameter and returns the three loss values and the (detached, :206)
hidden state.  There is no CPU path.
"""
import ctypes

import torch

from .. import _native
from ..options import TrainOptions
from ..predictor import WindowPredictor, _fptr


def _weights_struct(tensors):
    """hb_weights over a dict of CUDA tensors keyed like TransducerGRU.state_dict()."""
    w = _native.hb_weights()
    for layer, dst in (("gru_encoder", w.encoder), ("gru_decoder", w.decoder)):
        for d, rev in enumerate(("", "_reverse")):
            dst.weight_ih[d] = _fptr(tensors[f"{layer}.weight_ih_l0{rev}"])
            dst.weight_hh[d] = _fptr(tensors[f"{layer}.weight_hh_l0{rev}"])
            dst.bias_ih[d] = _fptr(tensors[f"{layer}.bias_ih_l0{rev}"])
            dst.bias_hh[d] = _fptr(tensors[f"{layer}.bias_hh_l0{rev}"])
    w.base_weight, w.base_bias = _fptr(tensors["dense1_base.weight"]), _fptr(tensors["dense1_base.bias"])
    w.rle_weight, w.rle_bias = _fptr(tensors["dense2_rle.weight"]), _fptr(tensors["dense2_rle.bias"])
    return w


class ChunkTrainer(object):
    def __init__(self, model, class_weights=None, device=0):
        if isinstance(device, torch.device):
            device = device.index or 0
        self.device = torch.device("cuda", int(device))
        self.model = model
        params = dict(model.named_parameters())
        for name, p in params.items():
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError(f"ChunkTrainer needs contiguous fp32 CUDA parameters ({name}); there is no CPU path")
            p.requires_grad_(True)
        self._params = params
        self._handle_owner = WindowPredictor(model.state_dict(), device=device, engine="fp32")
        self._lib = self._handle_owner._lib
        weights = TrainOptions.CLASS_WEIGHTS if class_weights is None else class_weights
        self.class_weights = torch.as_tensor(weights, dtype=torch.float32, device=self.device).contiguous()
        self._workspace = None
        self._loss = torch.zeros(3, dtype=torch.float32, device=self.device)

    def close(self):
        self._handle_owner.close()

    def step(self, image_chunk, hidden, label_base_chunk, label_rle_chunk, return_logits=False):
        """image_chunk float [B, W, F], hidden float [B, 2, H] or None, labels int64 [B, W] (CUDA)."""
        for name, t in (("image_chunk", image_chunk), ("label_base_chunk", label_base_chunk), ("label_rle_chunk", label_rle_chunk)):
            if not t.is_cuda:
                raise ValueError(f"{name} must be a CUDA tensor; there is no CPU path")
        x = image_chunk.to(torch.float32).contiguous()
        batch, width, features = x.shape
        if features != self._handle_owner.image_features:
            raise ValueError(f"image_chunk has {features} features, the model has {self._handle_owner.image_features}")
        lb = label_base_chunk.to(torch.int64).contiguous()
        lr = label_rle_chunk.to(torch.int64).contiguous()
        if lb.shape != (batch, width) or lr.shape != (batch, width):
            raise ValueError("labels must be [B, W]")
        h_in = None if hidden is None else hidden.to(device=self.device, dtype=torch.float32).contiguous()
        grads = {}
        for name, p in self._params.items():
            if p.grad is None or not p.grad.is_contiguous():
                p.grad = torch.zeros_like(p)
            grads[name] = p.grad
        w, g = _weights_struct({k: p.data for k, p in self._params.items()}), _weights_struct(grads)
        need = ctypes.c_size_t()
        handle = self._handle_owner._handle
        _native.check(self._lib.hb_train_workspace_bytes(handle, batch, width, ctypes.byref(need)))
        if self._workspace is None or self._workspace.numel() < need.value:
            self._workspace = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        h_out = torch.empty(batch, 2, self.model.hidden_size, dtype=torch.float32, device=self.device)
        base = torch.empty(batch, width, 5, dtype=torch.float32, device=self.device) if return_logits else None
        rle = torch.empty(batch, width, 11, dtype=torch.float32, device=self.device) if return_logits else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self._lib.hb_train_step_chunk(
            handle, ctypes.byref(w), ctypes.byref(g), x.data_ptr(), None if h_in is None else h_in.data_ptr(),
            lb.data_ptr(), lr.data_ptr(), self.class_weights.data_ptr(), batch, width, self._loss.data_ptr(), h_out.data_ptr(),
            None if base is None else base.data_ptr(), None if rle is None else rle.data_ptr(),
            self._workspace.data_ptr(), self._workspace.numel(), stream))
        loss, loss_base, loss_rle = self._loss.tolist()              # .item() in the reference loop (train.py:205-207)
        if loss != loss:
            # the loss kernel answers NaN for a label outside its class range (nn.CrossEntropyLoss asserts on the device)
            if int(lb.min()) < 0 or int(lb.max()) >= 5 or int(lr.min()) < 0 or int(lr.max()) >= 11:
                raise ValueError("labels out of range: base labels must be in [0, 5), run-length labels in [0, 11)")
        if return_logits:
            return loss, loss_base, loss_rle, h_out, base, rle
        return loss, loss_base, loss_rle, h_out
