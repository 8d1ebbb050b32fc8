"""ctypes binding of libethcnn_b200.so -- exactly the stub a maintainer of the reference would add
(see INTEGRATION.md).  Signatures follow include/ethcnn.h one to one."""
from __future__ import annotations

import ctypes as C
import math
import os
import sys
from typing import Optional, Tuple

import numpy as np

MODE_AI = 0
MODE_LDP = 1
PROBS_PER_CTU = 21
FC1_WIDTH = 448
STAGE_CONV, STAGE_FC1, STAGE_HEADS, STAGE_GATE = 0, 1, 2, 3
STAGE_NAMES = ("conv", "fc1", "heads", "gate")

Q_KERNEL_LAUNCHES, Q_N_DEVICES, Q_FC1_PATH, Q_TMA_LOADER_USED, Q_SM_COUNT, Q_CONV_PATH = 1, 2, 3, 4, 5, 6
OPT_FC1_PATH, OPT_CHUNK_CTUS, OPT_STAGED_OUTPUT, OPT_CONV_PATH = 1, 2, 3, 4
IPC_HANDLE_BYTES = 64

_LIB = None


class EthCnnError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("ethcnn error %d: %s" % (code, message))
        self.code = code


def library_path() -> str:
    return os.environ.get("ETHCNN_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libethcnn_b200.so")


def load_library() -> C.CDLL:
    """Load the CUDA extension; fails loudly (no fallback) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise EthCnnError(-2, "%s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C hevc-complexity-reduction_b200`; there is no CPU fallback" % path)
    lib = C.CDLL(path)
    vp, cp, i32, i64, sz = C.c_void_p, C.c_char_p, C.c_int, C.c_int64, C.c_size_t
    sig = {
        "ethcnn_abi_version": (i32, []),
        "ethcnn_last_error": (cp, []),
        "ethcnn_create": (i32, [cp, cp, i32, i32, C.POINTER(vp)]),
        "ethcnn_create_on_device": (i32, [cp, cp, i32, i32, C.POINTER(vp)]),
        "ethcnn_destroy": (None, [vp]),
        "ethcnn_predict_yuv_file": (i32, [vp, cp, i32, i32, i32, cp]),
        "ethcnn_predict_luma": (i32, [vp, vp, i32, i32, sz, i32, i32, vp]),
        "ethcnn_predict_luma_device": (i32, [vp, vp, i32, i32, sz, sz, i32, i32, vp, vp]),
        "ethcnn_predict_luma_map": (i32, [vp, vp, i32, i32, sz, i32, i32, vp, vp]),
        "ethcnn_predict_luma_device_map": (i32, [vp, vp, i32, i32, sz, sz, i32, i32, vp, vp, vp]),
        "ethcnn_set_decision_thresholds": (i32, [vp, vp]),
        "ethcnn_get_decision_thresholds": (i32, [vp, vp]),
        "ethcnn_export_fc1": (i32, [vp, vp, i32, i32, sz, i32, vp]),
        "ethcnn_decisions": (i32, [vp, vp, sz, vp, vp]),
        "ethcnn_ldp_step": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp]),
        "ethcnn_ldp_serve": (i32, [vp, cp, i32, i32]),
        "ethcnn_query": (i32, [vp, i32, C.POINTER(i64)]),
        "ethcnn_profile_enable": (i32, [vp, i32]),
        "ethcnn_profile_read": (i32, [vp, i32, C.POINTER(C.c_double), C.POINTER(i64), i32]),
        "ethcnn_set_option": (i32, [vp, i32, i64]),
        "ethcnn_peer_buffer_create": (i32, [vp, sz, C.POINTER(vp), vp]),
        "ethcnn_peer_buffer_open": (i32, [vp, vp, C.POINTER(vp)]),
        "ethcnn_peer_buffer_release": (i32, [vp, vp]),
        "ethcnn_serve": (i32, [vp, cp, i32, i32]),
        "ethcnn_predict_yuv_file_from": (i32, [vp, cp, cp, i32, i32, i32, cp]),
        "ethcnn_reload_thresholds": (i32, [vp, cp]),
        "ethcnn_request": (i32, [cp, cp, i32, i32, i32, cp]),
        "ethcnn_request_quit": (i32, [cp]),
        "ethcnn_request_error": (cp, []),
        "ethcnn_alloc_pinned": (vp, [sz]),
        "ethcnn_free_pinned": (None, [vp]),
        "ethcnn_debug_pack_model": (i32, [cp, C.c_float, vp, vp, vp, vp, vp, vp, vp]),
        "ethcnn_debug_pack_conv_tc": (i32, [cp, C.c_float, vp]),
        "ethcnn_debug_read_thresholds": (i32, [cp, vp]),
        "ethcnn_debug_f32_to_f16": (C.c_uint16, [C.c_float]),
        "ethcnn_debug_read_scratch": (i32, [vp, i32, sz, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise EthCnnError(rc, load_library().ethcnn_last_error().decode("utf-8", "replace"))


def ctu_grid(width: int, height: int) -> Tuple[int, int]:
    """(rows, cols) of 64x64 CTUs after the reference's zero padding (video_to_cu_depth.py:52-57)."""
    return math.ceil(height / 64), math.ceil(width / 64)


def unpack_decisions(dmap: np.ndarray) -> np.ndarray:
    """uint64 [n] decision words -> uint8 [n, 21] (2 = split only, 0 = no split, 1 = check both), entry k from bits [2k, 2k+1]."""
    dmap = np.ascontiguousarray(dmap, dtype=np.uint64).reshape(-1, 1)
    shifts = (2 * np.arange(PROBS_PER_CTU, dtype=np.uint64)).reshape(1, -1)
    return ((dmap >> shifts) & np.uint64(3)).astype(np.uint8)


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def request(socket_path: str, yuv_path: str, width: int, height: int, qp: int, out_path: str = "cu_depth.dat") -> bool:
    """Client of the resident server (include/ethcnn.h, ethcnn_request): True when the server did the work, False when
    nobody listens on socket_path (the caller then works in-process); raises EthCnnError when the server failed."""
    lib = load_library()
    rc = lib.ethcnn_request(os.fsencode(socket_path), os.fsencode(yuv_path), width, height, qp, os.fsencode(out_path))
    if rc == 0:
        return True
    msg = lib.ethcnn_request_error().decode("utf-8", "replace")
    if rc == -2 and "no server" in msg:
        return False
    if rc == -3 and "resident handle" in msg:   # the server holds other weights than this directory's checkpoint and refused
        sys.stderr.write("video_to_cu_depth: %s, working in-process\n" % msg)
        return False
    raise EthCnnError(rc, "server: " + msg)


def request_quit(socket_path: str) -> bool:
    return load_library().ethcnn_request_quit(os.fsencode(socket_path)) == 0


class EthCnn(object):
    """One predictor instance (weights resident on the device across calls).

    Mirrors the module-level state of video_to_cu_depth.py:14-29 (session + saver) and
    net_CNN.py:47 (thresholds read from Thr_info.txt)."""

    def __init__(self, model_dir: str = ".", thr_path: Optional[str] = None, mode: int = MODE_AI, n_gpus: int = 1,
                 device: Optional[int] = None):
        self._lib = load_library()
        self._h = C.c_void_p()
        md = os.fsencode(model_dir)
        tp = os.fsencode(thr_path) if thr_path is not None else None
        if device is None:
            _check(self._lib.ethcnn_create(md, tp, mode, n_gpus, C.byref(self._h)))
        else:
            _check(self._lib.ethcnn_create_on_device(md, tp, mode, device, C.byref(self._h)))
        self.mode = mode

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ethcnn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    # --- the reference's script as a call (video_to_cu_depth.py:120-145)
    def predict_yuv_file(self, yuv_path: str, width: int, height: int, qp: int, out_path: str = "cu_depth.dat") -> None:
        _check(self._lib.ethcnn_predict_yuv_file(self._h, os.fsencode(yuv_path), width, height, qp, os.fsencode(out_path)))

    # --- get_prob() for luma in host memory (video_to_cu_depth.py:75-118)
    def predict_yuv_file_from(self, client_dir: str, yuv_path: str, width: int, height: int, qp: int, out_path: str = "cu_depth.dat") -> None:
        """A resident handle answering like a fresh run of the reference script from `client_dir` (see include/ethcnn.h)."""
        _check(self._lib.ethcnn_predict_yuv_file_from(self._h, os.fsencode(client_dir), os.fsencode(yuv_path), width, height, qp,
                                                      os.fsencode(out_path)))

    def reload_thresholds(self, thr_path: Optional[str] = None) -> None:
        _check(self._lib.ethcnn_reload_thresholds(self._h, os.fsencode(thr_path) if thr_path else None))

    def predict_yuv_buffer(self, yuv: np.ndarray, width: int, height: int, qp: int) -> np.ndarray:
        """yuv: uint8 array holding whole 4:2:0 frames (Y, U, V planar).  Returns float32 [n_frames*nCTU, 21]."""
        yuv = np.ascontiguousarray(yuv, dtype=np.uint8).reshape(-1)
        frame_bytes = width * height * 3 // 2
        if yuv.size % frame_bytes != 0:
            raise EthCnnError(-1, "file_bytes % frame_bytes != 0")  # video_to_cu_depth.py:137
        return self.predict_luma(yuv, width, height, yuv.size // frame_bytes, qp, frame_stride=frame_bytes)

    def predict_luma(self, y: np.ndarray, width: int, height: int, n_frames: int, qp: int,
                     frame_stride: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        y = np.ascontiguousarray(y, dtype=np.uint8)
        if frame_stride is None:
            frame_stride = width * height
        need = (n_frames - 1) * frame_stride + width * height if n_frames > 0 else 0
        if y.size < need:
            raise EthCnnError(-1, "luma buffer too small")
        r, c = ctu_grid(width, height)
        if out is None:
            out = np.empty((n_frames * r * c, PROBS_PER_CTU), dtype=np.float32)
        _check(self._lib.ethcnn_predict_luma(self._h, _ptr(y), width, height, frame_stride, n_frames, qp, _ptr(out)))
        return out

    def predict_luma_ptr(self, y_ptr: int, width: int, height: int, frame_stride: int, n_frames: int, qp: int, out_ptr: int):
        """Raw host-pointer form (pinned buffers owned by the caller, e.g. torch pinned tensors)."""
        _check(self._lib.ethcnn_predict_luma(self._h, C.c_void_p(y_ptr), width, height, frame_stride, n_frames, qp,
                                             C.c_void_p(out_ptr)))

    def predict_luma_device(self, d_y: int, width: int, height: int, pitch: int, frame_stride: int, n_frames: int, qp: int,
                            d_out: int, stream: int = 0) -> None:
        """Device pointers (ints), asynchronous on `stream` (a cudaStream_t value)."""
        _check(self._lib.ethcnn_predict_luma_device(self._h, C.c_void_p(d_y), width, height, pitch, frame_stride, n_frames,
                                                    qp, C.c_void_p(d_out), C.c_void_p(stream)))

    # --- decision map (include/ethcnn.h: HM's threshold rule on the device, one 64-bit word of 21 two-bit decisions per CTU)
    def predict_luma_map(self, y: np.ndarray, width: int, height: int, n_frames: int, qp: int,
                         frame_stride: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
        """(float32 [n_frames*nCTU, 21], uint64 [n_frames*nCTU]) -- the rows of predict_luma plus the packed decisions."""
        y = np.ascontiguousarray(y, dtype=np.uint8)
        if frame_stride is None:
            frame_stride = width * height
        rows, cols = ctu_grid(width, height)
        out = np.empty((n_frames * rows * cols, PROBS_PER_CTU), dtype=np.float32)
        dmap = np.empty((n_frames * rows * cols,), dtype=np.uint64)
        _check(self._lib.ethcnn_predict_luma_map(self._h, _ptr(y), width, height, frame_stride, n_frames, qp, _ptr(out), _ptr(dmap)))
        return out, dmap

    def predict_luma_device_map(self, d_y: int, width: int, height: int, pitch: int, frame_stride: int, n_frames: int, qp: int,
                                d_out: int, d_map: int, stream: int = 0) -> None:
        _check(self._lib.ethcnn_predict_luma_device_map(self._h, C.c_void_p(d_y), width, height, pitch, frame_stride, n_frames,
                                                        qp, C.c_void_p(d_out), C.c_void_p(d_map), C.c_void_p(stream)))

    def set_decision_thresholds(self, thr6) -> None:
        thr = np.asarray(thr6, dtype=np.float32).reshape(6)
        _check(self._lib.ethcnn_set_decision_thresholds(self._h, _ptr(thr)))

    def get_decision_thresholds(self) -> np.ndarray:
        thr = np.empty(6, dtype=np.float32)
        _check(self._lib.ethcnn_get_decision_thresholds(self._h, _ptr(thr)))
        return thr

    def predict_ctus(self, ctus: np.ndarray, qp: int) -> np.ndarray:
        """ctus: uint8 [n, 64, 64].  Each CTU is treated as its own 64x64 frame, i.e. its own sub-batch of one
        (the reference run on a 64x64 'video'); returns float32 [n, 21]."""
        ctus = np.ascontiguousarray(ctus, dtype=np.uint8).reshape(-1, 64 * 64)
        return self.predict_luma(ctus, 64, 64, ctus.shape[0], qp)

    def export_fc1(self, y: np.ndarray, width: int, height: int, n_frames: int, frame_stride: Optional[int] = None) -> np.ndarray:
        """LDP FC1 tap (net_CNN_LSTM_one_step.py:151-199): float32 [n_frames*nCTU, 448]."""
        y = np.ascontiguousarray(y, dtype=np.uint8)
        if frame_stride is None:
            frame_stride = width * height
        r, c = ctu_grid(width, height)
        out = np.empty((n_frames * r * c, FC1_WIDTH), dtype=np.float32)
        _check(self._lib.ethcnn_export_fc1(self._h, _ptr(y), width, height, frame_stride, n_frames, _ptr(out)))
        return out

    def ldp_step(self, luma: np.ndarray, qp: int, i_frame: int, state_in: Optional[np.ndarray] = None):
        """One frame of the deployed LDP predictor (resi_to_cu_depth_LDP.py:114-129): luma uint8 [H, W] of the residue
        frame, state_in float32 [nCTU, 1, 2, 448] or None (zeros).  Returns (prob [nCTU, 21], state_out [nCTU, 1, 2, 448])."""
        luma = np.ascontiguousarray(luma, dtype=np.uint8)
        h, w = luma.shape
        r, c = ctu_grid(w, h)
        n = r * c
        prob = np.empty((n, PROBS_PER_CTU), dtype=np.float32)
        state_out = np.empty((n, 1, 2, FC1_WIDTH), dtype=np.float32)
        sin = None
        if state_in is not None:
            sin = np.ascontiguousarray(state_in, dtype=np.float32)
            if sin.size != n * 2 * FC1_WIDTH:
                raise EthCnnError(-1, "state_in has the wrong size")
        _check(self._lib.ethcnn_ldp_step(self._h, _ptr(luma), w, h, qp, i_frame, _ptr(sin) if sin is not None else None,
                                         _ptr(state_out), _ptr(prob)))
        return prob, state_out

    def ldp_serve(self, directory: str = ".", max_frames: int = 0, idle_timeout_ms: int = 0) -> int:
        """The file-signal daemon loop (README.md:64-84); returns the number of frames served."""
        rc = self._lib.ethcnn_ldp_serve(self._h, os.fsencode(directory), max_frames, idle_timeout_ms)
        if rc < 0:
            _check(rc)
        return rc

    def serve(self, socket_path: str, max_requests: int = 0, idle_timeout_ms: int = 0) -> int:
        """The resident All-Intra server loop (include/ethcnn.h, ethcnn_serve); returns the number of requests served."""
        rc = self._lib.ethcnn_serve(self._h, os.fsencode(socket_path), max_requests, idle_timeout_ms)
        if rc < 0:
            raise EthCnnError(rc, "cannot serve on %s" % socket_path)
        return rc

    def decisions(self, prob: np.ndarray, thr6=(0.5,) * 6) -> np.ndarray:
        """HM's threshold rule on the device (TEncCu.cpp:448-462): uint8 array shaped like prob."""
        prob = np.ascontiguousarray(prob, dtype=np.float32).reshape(-1, PROBS_PER_CTU)
        thr = np.asarray(thr6, dtype=np.float32)
        out = np.empty(prob.shape, dtype=np.uint8)
        _check(self._lib.ethcnn_decisions(self._h, _ptr(prob), prob.shape[0], _ptr(thr), _ptr(out)))
        return out

    def debug_read_scratch(self, what: int, n_ctus: int) -> np.ndarray:
        """Testing hook: intermediates of the last chunk (0 = conv features [n,2688], 1 = FC1 [n,448])."""
        out = np.empty((n_ctus, 2688 if what == 0 else FC1_WIDTH), dtype=np.float32)
        _check(self._lib.ethcnn_debug_read_scratch(self._h, what, n_ctus, _ptr(out)))
        return out

    # --- introspection / tuning
    def query(self, what: int) -> int:
        v = C.c_int64()
        _check(self._lib.ethcnn_query(self._h, what, C.byref(v)))
        return int(v.value)

    @property
    def kernel_launches(self) -> int:
        return self.query(Q_KERNEL_LAUNCHES)

    def set_option(self, option: int, value: int) -> None:
        _check(self._lib.ethcnn_set_option(self._h, option, value))

    # --- gather buffers in peer memory (include/ethcnn.h: "Multi-GPU gather without a collective")
    def peer_buffer_create(self, n_bytes: int) -> Tuple[int, bytes]:
        """(device pointer, 64-byte IPC handle) of a zero-filled buffer on this handle's device."""
        p = C.c_void_p()
        hb = (C.c_uint8 * IPC_HANDLE_BYTES)()
        _check(self._lib.ethcnn_peer_buffer_create(self._h, n_bytes, C.byref(p), C.cast(hb, C.c_void_p)))
        return int(p.value), bytes(hb)

    def peer_buffer_open(self, handle: bytes) -> int:
        """Map a buffer another process exported; raises EthCnnError when the devices cannot reach each other."""
        if len(handle) != IPC_HANDLE_BYTES:
            raise EthCnnError(-1, "an IPC handle has %d bytes" % IPC_HANDLE_BYTES)
        p = C.c_void_p()
        hb = (C.c_uint8 * IPC_HANDLE_BYTES).from_buffer_copy(handle)
        _check(self._lib.ethcnn_peer_buffer_open(self._h, C.cast(hb, C.c_void_p), C.byref(p)))
        return int(p.value)

    def peer_buffer_release(self, d_ptr: int) -> None:
        _check(self._lib.ethcnn_peer_buffer_release(self._h, C.c_void_p(d_ptr)))

    def profile_enable(self, on: bool = True) -> None:
        _check(self._lib.ethcnn_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, stage: int, reset: bool = False) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_int64()
        _check(self._lib.ethcnn_profile_read(self._h, stage, C.byref(ms), C.byref(n), 1 if reset else 0))
        return float(ms.value), int(n.value)
