"""ctypes view of oracle/_ref/libyolo2ref_{scalar,fast}.so -- the UNMODIFIED reference CPU path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs, never by the product package (yolo2_light_b200/).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

# reference enums (additionally.h:68-70, :376-403)
LAYER_TYPES = ["CONVOLUTIONAL", "DECONVOLUTIONAL", "CONNECTED", "MAXPOOL", "SOFTMAX", "DETECTION", "DROPOUT", "CROP",
               "ROUTE", "COST", "NORMALIZATION", "AVGPOOL", "LOCAL", "SHORTCUT", "ACTIVE", "RNN", "GRU", "CRNN",
               "BATCHNORM", "NETWORK", "XNOR", "REGION", "YOLO", "UPSAMPLE", "REORG", "BLANK"]
ACTIVATIONS = ["LOGISTIC", "RELU", "RELIE", "LINEAR", "RAMP", "TANH", "PLSE", "LEAKY", "ELU", "LOGGY", "STAIR",
               "HARDTAN", "LHTAN", "SELU"]
INT_FIELDS = ["type", "activation", "batch_normalize", "batch", "h", "w", "c", "n", "size", "stride", "pad",
              "out_h", "out_w", "out_c", "inputs", "outputs", "xnor", "binary", "quantized", "index",
              "classes", "coords", "softmax", "total", "reverse", "lda_align", "new_lda", "bit_align",
              "align_bit_weights_size", "dontload", "dontloadscales", "use_bin_output", "max_boxes", "groups"]


def lib_path(kind: str = "scalar") -> str:
    return os.path.join(REF_DIR, f"libyolo2ref_{kind}.so")


def available(kind: str = "scalar") -> bool:
    return os.path.exists(lib_path(kind))


_libs = {}


def _load(kind: str):
    if kind in _libs:
        return _libs[kind]
    lib = C.CDLL(lib_path(kind), mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
    lib.refh_create.restype = C.c_void_p
    lib.refh_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.refh_num_layers.argtypes = [C.c_void_p]
    lib.refh_net_ints.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.refh_net_input_calibration.restype = C.POINTER(C.c_float)
    lib.refh_net_input_calibration.argtypes = [C.c_void_p]
    lib.refh_layer_ints.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.refh_layer_floats.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    lib.refh_layer_ptr.restype = C.c_void_p
    lib.refh_layer_ptr.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.refh_predict.restype = C.POINTER(C.c_float)
    lib.refh_predict.argtypes = [C.c_void_p, C.c_void_p]
    lib.refh_forward_layer.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.refh_time_predict.restype = C.c_double
    lib.refh_time_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.refh_get_boxes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                   C.POINTER(C.c_int)]
    lib.refh_set_quiet.argtypes = [C.c_int]
    lib.refh_load_resize_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    if hasattr(lib, "refh_predict_b200"):
        lib.refh_predict_b200.restype = C.POINTER(C.c_float)
        lib.refh_predict_b200.argtypes = [C.c_void_p, C.c_void_p]
    _libs[kind] = lib
    return lib


class RefNet:
    """The reference's ``network`` after the main.c:160-171 preparation sequence."""

    def __init__(self, cfg: str, weights: Optional[str], batch: int = 1, quantized: int = 0, prep: int = 7,
                 kind: str = "scalar"):
        self.lib = _load(kind)
        self.h = self.lib.refh_create(cfg.encode(), (weights or "").encode(), batch, quantized, prep)
        ints = (C.c_int * 16)()
        self.lib.refh_net_ints(self.h, ints)
        self.n, self.batch, self.height, self.width, self.channels = ints[0], ints[1], ints[2], ints[3], ints[4]
        self.inputs, self.outputs, self.input_calibration_size = ints[5], ints[6], ints[7]
        self.quantized = quantized
        self.layers = [self._layer(i) for i in range(self.n)]

    def _layer(self, i: int) -> dict:
        ints = (C.c_int * 40)()
        self.lib.refh_layer_ints(self.h, i, ints)
        d = {k: ints[j] for j, k in enumerate(INT_FIELDS)}
        d["type_name"] = LAYER_TYPES[d["type"]]
        d["activation_name"] = ACTIVATIONS[d["activation"]]
        fl = (C.c_float * 8)()
        self.lib.refh_layer_floats(self.h, i, fl)
        d.update(weights_quant_multipler=fl[0], input_quant_multipler=fl[1], output_multipler=fl[2],
                 scale=fl[3], bflops=fl[4])
        return d

    def input_calibration(self) -> np.ndarray:
        p = self.lib.refh_net_input_calibration(self.h)
        if not p or self.input_calibration_size == 0:
            return np.zeros(0, np.float32)
        return np.ctypeslib.as_array(p, shape=(self.input_calibration_size,)).copy()

    def array(self, i: int, what: str, count: int, dtype=np.float32) -> Optional[np.ndarray]:
        p = self.lib.refh_layer_ptr(self.h, i, what.encode())
        if not p:
            return None
        buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=count).copy()

    def output(self, i: int) -> np.ndarray:
        L = self.layers[i]
        a = self.array(i, "output", L["outputs"] * L["batch"])
        if L["type_name"] in ("CONVOLUTIONAL", "MAXPOOL", "ROUTE", "UPSAMPLE", "SHORTCUT", "REORG", "YOLO"):
            return a.reshape(L["batch"], L["out_c"], L["out_h"], L["out_w"])
        return a.reshape(L["batch"], -1)

    def set_output(self, i: int, a: np.ndarray) -> None:
        """Plant an activation in the reference layer's host l.output (what ROUTE / SHORTCUT layers read from their sources)."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        L = self.layers[i]
        assert a.size == L["outputs"] * L["batch"], (a.shape, L["outputs"], L["batch"])
        p = self.lib.refh_layer_ptr(self.h, i, b"output")
        C.memmove(p, a.ctypes.data, a.nbytes)

    def predict(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.size == self.inputs * self.batch, (x.shape, self.inputs, self.batch)
        self.lib.refh_predict(self.h, x.ctypes.data_as(C.c_void_p))
        return self.output(self.n - 1)

    def predict_b200(self, x: np.ndarray) -> np.ndarray:
        """network_predict_b200 of integration/yolo2_light_b200_glue.c (kind='dropin' only)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        self.lib.refh_predict_b200(self.h, x.ctypes.data_as(C.c_void_p))
        return self.output(self.n - 1)

    def forward_conv_b200(self, i: int, x: np.ndarray, use_q_rule: Optional[bool] = None) -> np.ndarray:
        """forward_convolutional_layer_b200[_q](layer l, network_state state) of the glue (kind='dropin'): the reference's
        by-value per-layer call shape served by the engine; the result lands in the reference layer's host l.output."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        q = self.quantized if use_q_rule is None else int(use_q_rule)
        self.lib.refh_forward_conv_b200.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self.lib.refh_forward_conv_b200(self.h, i, x.ctypes.data_as(C.c_void_p), q)
        return self.output(i)

    def predict_b200_batch(self, x: np.ndarray, ngpus: int) -> np.ndarray:
        """network_predict_b200_batch of the glue: x = float32[nimg, c, h, w]; returns the last layer [nimg, outputs]."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        nimg = x.shape[0]
        self.lib.refh_predict_b200_batch.restype = C.POINTER(C.c_float)
        self.lib.refh_predict_b200_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        p = self.lib.refh_predict_b200_batch(self.h, x.ctypes.data_as(C.c_void_p), nimg, ngpus)
        return np.ctypeslib.as_array(p, shape=(nimg, self.layers[-1]["outputs"])).copy()

    def time_predict_b200(self, x: np.ndarray, reps: int, decode: bool = False, thresh: float = 0.24, nms: float = 0.45) -> float:
        x = np.ascontiguousarray(x, dtype=np.float32)
        self.lib.refh_time_predict_b200.restype = C.c_double
        self.lib.refh_time_predict_b200.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]
        return float(self.lib.refh_time_predict_b200(self.h, x.ctypes.data_as(C.c_void_p), reps, int(decode), thresh, nms))

    def forward_layer(self, i: int, x: np.ndarray, use_q_rule: Optional[bool] = None) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        q = self.quantized if use_q_rule is None else int(use_q_rule)
        self.lib.refh_forward_layer(self.h, i, x.ctypes.data_as(C.c_void_p), q)
        return self.output(i)

    def time_predict(self, x: np.ndarray, reps: int = 1) -> float:
        x = np.ascontiguousarray(x, dtype=np.float32)
        return float(self.lib.refh_time_predict(self.h, x.ctypes.data_as(C.c_void_p), reps))

    def get_boxes_b200(self, w: int, h: int, thresh: float, nms: float, max_out: int = 8192):
        """get_network_boxes_nms_b200 of the glue (kind='dropin', after predict_b200): device-side decode + NMS."""
        classes = self.layers[-1]["classes"]
        out = np.zeros((max_out, 6 + classes), np.float32)
        self.lib.refh_get_boxes_b200.restype = C.c_int
        self.lib.refh_get_boxes_b200.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int]
        n = self.lib.refh_get_boxes_b200(self.h, w, h, thresh, nms, out.ctypes.data_as(C.c_void_p), max_out)
        return out[:min(n, max_out)]

    def get_boxes(self, w: int, h: int, thresh: float, nms: float, max_out: int = 200000):
        classes = self.layers[-1]["classes"]
        out = np.zeros((max_out, 6 + classes), np.float32)
        cl = C.c_int()
        n = self.lib.refh_get_boxes(self.h, w, h, thresh, nms, out.ctypes.data_as(C.c_void_p), max_out, C.byref(cl))
        return out[:min(n, max_out)]


def validate_map(datacfg: str, cfg: str, weights: str, thresh_calc_avg_iou: float, quantized: int, iou_thresh: float,
                 out_path: str, kind: str = "scalar") -> str:
    """Runs the reference's validate_detector_map and returns what it printed."""
    lib = _load(kind)
    lib.refh_validate_map.restype = None
    lib.refh_validate_map.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_float, C.c_int, C.c_float, C.c_char_p]
    lib.refh_validate_map(datacfg.encode(), cfg.encode(), weights.encode(), thresh_calc_avg_iou, quantized, iou_thresh,
                          out_path.encode())
    return open(out_path).read()


def entropy_calibration(src: np.ndarray, bin_width: float = 1.0 / 16, max_bin: int = 4096, kind: str = "scalar") -> float:
    """The reference's entropy_calibration on one float array (prints one line to stdout, like the reference)."""
    lib = _load(kind)
    a = np.ascontiguousarray(src, dtype=np.float32).ravel()
    lib.refh_entropy_calibration.restype = C.c_float
    lib.refh_entropy_calibration.argtypes = [C.c_void_p, C.c_long, C.c_float, C.c_int]
    return float(lib.refh_entropy_calibration(a.ctypes.data_as(C.c_void_p), a.size, bin_width, max_bin))


def load_resize_u8(img_hwc: np.ndarray, out_w: int, out_h: int, kind: str = "scalar") -> np.ndarray:
    """The reference's load_image_stb conversion + resize_image on one u8 HWC image -> float32[c, out_h, out_w]."""
    lib = _load(kind)
    img = np.ascontiguousarray(img_hwc, dtype=np.uint8)
    h, w, c = img.shape
    out = np.empty((c, out_h, out_w), np.float32)
    lib.refh_load_resize_u8(img.ctypes.data_as(C.c_void_p), w, h, c, out_w, out_h, out.ctypes.data_as(C.c_void_p))
    return out
