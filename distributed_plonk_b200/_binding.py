"""ctypes binding of the C ABI declared in include/dplonk.h (one Python method per entry point)."""
from __future__ import annotations

import ctypes as C

import numpy as np

DP_OK, DP_E_ARG, DP_E_STATE, DP_E_OOM, DP_E_CUDA, DP_E_COMM = 0, -1, -2, -3, -4, -5
FR_BYTES, G1_AFFINE_BYTES, G1_PROJECTIVE_BYTES = 32, 104, 144

EXPORTS = [
    "dp_create", "dp_destroy", "dp_last_error", "dp_version", "dp_init", "dp_msm", "dp_commit", "dp_fft_init",
    "dp_fft1", "dp_fft1_rows", "dp_fft2_prepare", "dp_fft_exchange_begin", "dp_fft_exchange_end", "dp_fft2",
    "dp_ntt", "dp_round1", "dp_get_wire", "dp_peer_arena_create", "dp_peer_attach", "dp_last_timing",
    "dp_launch_count", "dp_sync", "dp_msm_dev", "dp_ntt_dev", "dp_fft_dev", "dp_debug_set_limits",
    "dp_last_msm_breakdown", "dp_msm_tuning", "dp_msm_tuning_all", "dp_debug_gen_bases", "dp_fft_dev_rows", "dp_fft_dev_cols", "dp_peer_ready", "dp_fft_dev_rows_p2p", "dp_fft_dev_p2p", "dp_msm_dev_batch", "dp_perm_product", "dp_msm_batch", "dp_perm_product_dev",
    "dp_quotient_evals", "dp_quotient_evals_dev", "dp_poly_eval", "dp_poly_eval_dev", "dp_poly_lincomb", "dp_poly_lincomb_dev",
    "dp_poly_div_linear", "dp_poly_div_linear_dev", "dp_init_compressed", "dp_get_bases",
    "dp_msm_submit", "dp_msm_collect", "dp_poly_put", "dp_poly_ptr", "dp_poly_get", "dp_poly_free", "dp_commit_dev",
    "dp_fft_exchange_begin_async", "dp_compute_stream", "dp_fft_dev_p2p_async", "dp_fft1_rows_short", "dp_fft_dev_hint_valid_cols", "dp_ntt_dev_padded", "dp_debug_set_three_pass",
]


class DpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dplonk error {code}: {msg}")
        self.code = code


class FftWorkload(C.Structure):
    """utils.rs:3-19 / hello_world.capnp:8-13"""
    _fields_ = [("row_start", C.c_uint64), ("row_end", C.c_uint64), ("col_start", C.c_uint64), ("col_end", C.c_uint64)]


class QuotientArgs(C.Structure):
    """dp_quotient_args (include/dplonk.h): 25 polynomial-sized arrays + the challenges"""
    _fields_ = [("selectors", C.c_void_p * 13), ("sigmas", C.c_void_p * 5), ("wires", C.c_void_p * 5), ("perm", C.c_void_p),
                ("pub_input", C.c_void_p), ("k", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p), ("gamma", C.c_void_p)]


def bind(cdll: C.CDLL) -> C.CDLL:
    u64, vp, i, sz, u32 = C.c_uint64, C.c_void_p, C.c_int, C.c_size_t, C.c_uint32
    sig = {
        "dp_create": (i, [i, u64, u64, C.POINTER(vp)]),
        "dp_destroy": (i, [vp]),
        "dp_last_error": (C.c_char_p, [vp]),
        "dp_version": (C.c_char_p, []),
        "dp_init": (i, [vp, vp, sz, u64, u64]),
        "dp_msm": (i, [vp, u64, u64, vp, sz, vp]),
        "dp_commit": (i, [vp, vp, sz, vp]),
        "dp_fft_init": (i, [vp, u64, C.POINTER(FftWorkload), sz, i, i, i]),
        "dp_fft1": (i, [vp, u64, u64, vp, sz]),
        "dp_fft1_rows": (i, [vp, u64, u64, u64, vp]),
        "dp_fft1_rows_short": (i, [vp, u64, u64, u64, vp, sz]),
        "dp_fft2_prepare": (i, [vp, u64]),
        "dp_fft_exchange_begin": (i, [vp, u64, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "dp_fft_exchange_end": (i, [vp, u64]),
        "dp_fft_exchange_begin_async": (i, [vp, u64, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "dp_compute_stream": (i, [vp, C.POINTER(vp)]),
        "dp_fft2": (i, [vp, u64, vp, sz]),
        "dp_ntt": (i, [vp, vp, sz, u32, i, i]),
        "dp_round1": (i, [vp, vp, sz, vp, vp]),
        "dp_get_wire": (i, [vp, vp, sz, C.POINTER(sz)]),
        "dp_peer_arena_create": (i, [vp, u64, vp]),
        "dp_peer_attach": (i, [vp, u64, vp]),
        "dp_last_timing": (i, [vp, C.POINTER(C.c_float), C.POINTER(u64)]),
        "dp_launch_count": (u64, [vp]),
        "dp_sync": (i, [vp]),
        "dp_msm_dev": (i, [vp, u64, u64, vp, sz, vp]),
        "dp_ntt_dev": (i, [vp, vp, u32, i, i]),
        "dp_ntt_dev_padded": (i, [vp, vp, sz, C.c_uint32, i, i, i]),
        "dp_fft_dev": (i, [vp, vp, vp, i, i, i]),
        "dp_debug_set_limits": (i, [vp, u32, u32, i]),
        "dp_last_msm_breakdown": (i, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "dp_msm_tuning": (i, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "dp_msm_tuning_all": (i, [vp, C.POINTER(C.c_float)]),
        "dp_debug_gen_bases": (i, [vp, u64, sz, vp]),
        "dp_fft_dev_rows": (i, [vp, vp, i, i, i, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "dp_fft_dev_cols": (i, [vp, vp]),
        "dp_peer_ready": (i, [vp]),
        "dp_perm_product": (i, [vp, vp, vp, vp, sz, sz, vp, vp, vp]),
        "dp_perm_product_dev": (i, [vp, vp, vp, vp, sz, sz, vp, vp, vp]),
        "dp_msm_batch": (i, [vp, sz, C.POINTER(u64), C.POINTER(u64), C.POINTER(vp), C.POINTER(sz), C.POINTER(vp)]),
        "dp_msm_dev_batch": (i, [vp, sz, C.POINTER(u64), C.POINTER(u64), C.POINTER(vp), C.POINTER(sz), C.POINTER(vp)]),
        "dp_fft_dev_rows_p2p": (i, [vp, vp, i, i, i]),
        "dp_fft_dev_p2p": (i, [vp, vp, vp, i, i, i]),
        "dp_fft_dev_p2p_async": (i, [vp, vp, vp, i, i, i]),
        "dp_fft_dev_hint_valid_cols": (i, [vp, i, u64]),
        "dp_debug_set_three_pass": (i, [vp, C.c_uint32]),
        "dp_poly_put": (i, [vp, u64, vp, sz, sz]),
        "dp_poly_ptr": (i, [vp, u64, C.POINTER(vp), C.POINTER(sz)]),
        "dp_poly_get": (i, [vp, u64, sz, sz, vp]),
        "dp_poly_free": (i, [vp, u64]),
        "dp_commit_dev": (i, [vp, vp, sz, vp]),
        "dp_msm_submit": (i, [vp, u64, u64, u64, vp, sz]),
        "dp_msm_collect": (i, [vp, u64, vp]),
        "dp_init_compressed": (i, [vp, vp, sz, u64, u64, i]),
        "dp_get_bases": (i, [vp, u64, sz, vp]),
        "dp_quotient_evals": (i, [vp, C.POINTER(QuotientArgs), vp]),
        "dp_quotient_evals_dev": (i, [vp, C.POINTER(QuotientArgs), vp]),
        "dp_poly_eval": (i, [vp, vp, sz, vp, vp]),
        "dp_poly_eval_dev": (i, [vp, vp, sz, vp, vp]),
        "dp_poly_lincomb": (i, [vp, C.POINTER(vp), C.POINTER(sz), vp, sz, vp, sz]),
        "dp_poly_lincomb_dev": (i, [vp, C.POINTER(vp), C.POINTER(sz), vp, sz, vp, sz]),
        "dp_poly_div_linear": (i, [vp, vp, sz, vp, vp, vp]),
        "dp_poly_div_linear_dev": (i, [vp, vp, sz, vp, vp, vp]),
    }
    assert set(sig) == set(EXPORTS)
    for name, (res, args) in sig.items():
        f = getattr(cdll, name)  # AttributeError here = the library does not export what dplonk.h declares
        f.restype = res
        f.argtypes = args
    return cdll


def _addr(buf) -> int:
    """host address of a numpy array / bytes-like / raw int address."""
    if isinstance(buf, int):
        return buf
    if isinstance(buf, np.ndarray):
        assert buf.flags["C_CONTIGUOUS"]
        return buf.ctypes.data
    return C.addressof(C.c_char.from_buffer(buf))


class Context:
    """One dp_ctx (one GPU).  Thin: arguments are numpy arrays in the reference's raw layouts."""

    def __init__(self, cdll: C.CDLL, device: int = 0, me: int = 0, n_workers: int = 1):
        self.lib = cdll
        self.me, self.n_workers = me, n_workers
        h = C.c_void_p()
        rc = cdll.dp_create(device, me, n_workers, C.byref(h))
        if rc != DP_OK:
            raise DpError(rc, (cdll.dp_last_error(None) or b"").decode())
        self.h = h

    def close(self):
        if self.h:
            self.lib.dp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != DP_OK:
            raise DpError(rc, (self.lib.dp_last_error(self.h) or b"").decode())

    # ---- PlonkSlave surface
    def init(self, bases: np.ndarray, domain_size: int, quot_domain_size: int):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        n = bases.size // G1_AFFINE_BYTES
        self._ck(self.lib.dp_init(self.h, _addr(bases) if n else None, n, domain_size, quot_domain_size))

    def msm(self, start: int, end: int, scalars: np.ndarray) -> np.ndarray:
        scalars = np.ascontiguousarray(scalars)
        n = scalars.nbytes // 32
        out = np.zeros(G1_PROJECTIVE_BYTES, dtype=np.uint8)
        self._ck(self.lib.dp_msm(self.h, start, end, _addr(scalars) if n else None, n, _addr(out)))
        return out

    def commit(self, coeffs: np.ndarray) -> np.ndarray:
        coeffs = np.ascontiguousarray(coeffs)
        n = coeffs.nbytes // 32
        out = np.zeros(G1_PROJECTIVE_BYTES, dtype=np.uint8)
        self._ck(self.lib.dp_commit(self.h, _addr(coeffs) if n else None, n, _addr(out)))
        return out

    def fft_init(self, task_id: int, workloads, is_quot: bool, is_inv: bool, is_coset: bool):
        arr = (FftWorkload * len(workloads))(*[FftWorkload(*w) for w in workloads])
        self._ck(self.lib.dp_fft_init(self.h, task_id, arr, len(workloads), int(is_quot), int(is_inv), int(is_coset)))

    def fft1(self, task_id: int, i: int, row: np.ndarray):
        row = np.ascontiguousarray(row)
        self._ck(self.lib.dp_fft1(self.h, task_id, i, _addr(row), row.nbytes // 32))

    def fft1_rows(self, task_id: int, i_first: int, rows: np.ndarray, n_rows: int):
        rows = np.ascontiguousarray(rows)
        self._ck(self.lib.dp_fft1_rows(self.h, task_id, i_first, n_rows, _addr(rows)))

    def fft1_rows_short(self, task_id: int, i_first: int, rows, n_rows: int, row_len: int):
        """n_rows rows of row_len leading entries each (compact array or host address); the tails are implicit zeros"""
        self._ck(self.lib.dp_fft1_rows_short(self.h, task_id, i_first, n_rows, _addr(rows), row_len))

    def fft2_prepare(self, task_id: int):
        self._ck(self.lib.dp_fft2_prepare(self.h, task_id))

    def fft_exchange_begin(self, task_id: int):
        s, r, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._ck(self.lib.dp_fft_exchange_begin(self.h, task_id, C.byref(s), C.byref(r), C.byref(n)))
        return s.value, r.value, n.value

    def fft_exchange_begin_async(self, task_id: int):
        """like fft_exchange_begin, without waiting: the buffers are valid for work on compute_stream()"""
        s, r, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._ck(self.lib.dp_fft_exchange_begin_async(self.h, task_id, C.byref(s), C.byref(r), C.byref(n)))
        return s.value, r.value, n.value

    def compute_stream(self) -> int:
        st = C.c_void_p()
        self._ck(self.lib.dp_compute_stream(self.h, C.byref(st)))
        return st.value or 0

    def fft_exchange_end(self, task_id: int):
        self._ck(self.lib.dp_fft_exchange_end(self.h, task_id))

    def fft2(self, task_id: int, n_cols: int, r: int) -> np.ndarray:
        out = np.empty((n_cols, r, 4), dtype=np.uint64)
        self._ck(self.lib.dp_fft2(self.h, task_id, _addr(out), out.nbytes))
        return out

    def ntt(self, data: np.ndarray, log_n: int, is_inv: bool, is_coset: bool) -> np.ndarray:
        n = data.nbytes // 32
        buf = np.zeros((1 << log_n, 4), dtype=np.uint64)
        buf.reshape(-1)[: n * 4] = np.ascontiguousarray(data).view(np.uint64).reshape(-1)
        self._ck(self.lib.dp_ntt(self.h, _addr(buf), n, log_n, int(is_inv), int(is_coset)))
        return buf

    def round1(self, evals: np.ndarray, blind: np.ndarray | None) -> np.ndarray:
        evals = np.ascontiguousarray(evals)
        out = np.zeros(G1_PROJECTIVE_BYTES, dtype=np.uint8)
        b = np.ascontiguousarray(blind) if blind is not None else None
        self._ck(self.lib.dp_round1(self.h, _addr(evals), evals.nbytes // 32, _addr(b) if b is not None else None, _addr(out)))
        return out

    def perm_product(self, wires: np.ndarray, id_perm: np.ndarray, sigma_perm: np.ndarray, beta: np.ndarray, gamma: np.ndarray) -> np.ndarray:
        """round-2 grand product (dispatcher2.rs:329-345); inputs [n_types, n, 4] u64 raw Fr"""
        a = [np.ascontiguousarray(x, dtype=np.uint64) for x in (wires, id_perm, sigma_perm, beta, gamma)]
        n_types, n = a[0].shape[0], a[0].shape[1]
        out = np.empty((n, 4), dtype=np.uint64)
        self._ck(self.lib.dp_perm_product(self.h, _addr(a[0]), _addr(a[1]), _addr(a[2]), n_types, n, _addr(a[3]), _addr(a[4]), _addr(out)))
        return out

    def perm_product_dev(self, wires_ptr: int, id_ptr: int, sigma_ptr: int, n_types: int, n: int, beta: np.ndarray, gamma: np.ndarray, out_ptr: int):
        b, g = np.ascontiguousarray(beta, dtype=np.uint64), np.ascontiguousarray(gamma, dtype=np.uint64)
        self._ck(self.lib.dp_perm_product_dev(self.h, wires_ptr, id_ptr, sigma_ptr, n_types, n, _addr(b), _addr(g), out_ptr))

    # ---- worker-resident polynomials
    def poly_put(self, poly_id: int, coeffs: np.ndarray, capacity: int = 0) -> int:
        """store [n,4] raw Fr under poly_id (zero-extended to `capacity`); returns the device address"""
        a = np.ascontiguousarray(coeffs, dtype=np.uint64)
        self._ck(self.lib.dp_poly_put(self.h, poly_id, _addr(a) if a.size else None, a.size // 4, capacity))
        return self.poly_ptr(poly_id)[0]

    def poly_ptr(self, poly_id: int):
        d, cap = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.dp_poly_ptr(self.h, poly_id, C.byref(d), C.byref(cap)))
        return d.value, cap.value

    def poly_get(self, poly_id: int, n: int | None = None, offset: int = 0) -> np.ndarray:
        if n is None:
            n = self.poly_ptr(poly_id)[1] - offset
        out = np.empty((n, 4), dtype=np.uint64)
        self._ck(self.lib.dp_poly_get(self.h, poly_id, offset, n, _addr(out) if n else None))
        return out

    def poly_free(self, poly_id: int):
        self._ck(self.lib.dp_poly_free(self.h, poly_id))

    def commit_dev(self, coeffs_ptr: int, n: int) -> np.ndarray:
        out = np.zeros(G1_PROJECTIVE_BYTES, dtype=np.uint8)
        self._ck(self.lib.dp_commit_dev(self.h, coeffs_ptr, n, _addr(out)))
        return out

    def msm_submit(self, job_id: int, start: int, end: int, scalars, n: int | None = None):
        """asynchronous varMsm: scalars = [n,4] host array (kept alive by the caller) or a host pointer + n"""
        if isinstance(scalars, int):
            self._ck(self.lib.dp_msm_submit(self.h, job_id, start, end, scalars, n))
        else:
            a = np.ascontiguousarray(scalars)
            self._keep = getattr(self, "_keep", {})
            self._keep[job_id] = a
            self._ck(self.lib.dp_msm_submit(self.h, job_id, start, end, _addr(a) if a.size else None, a.nbytes // 32))

    def msm_collect(self, job_id: int) -> np.ndarray:
        out = np.zeros(G1_PROJECTIVE_BYTES, dtype=np.uint8)
        try:
            self._ck(self.lib.dp_msm_collect(self.h, job_id, _addr(out)))
        finally:
            getattr(self, "_keep", {}).pop(job_id, None)
        return out

    def init_compressed(self, bases48: np.ndarray, domain_size: int, quot_domain_size: int, check_subgroup: bool = True):
        """dp_init from ark-serialize compressed points ([n, 48] uint8)"""
        b = np.ascontiguousarray(bases48, dtype=np.uint8)
        self._ck(self.lib.dp_init_compressed(self.h, _addr(b) if b.size else None, b.size // 48, domain_size, quot_domain_size, int(check_subgroup)))

    def get_bases(self, start: int, n: int) -> np.ndarray:
        out = np.zeros((n, 104), dtype=np.uint8)
        self._ck(self.lib.dp_get_bases(self.h, start, n, _addr(out) if n else None))
        return out

    # ---- rounds 3-5 ("next" row 1): plain = host arrays, *_dev = device pointers (ints)
    @staticmethod
    def _quotient_args(selectors, sigmas, wires, perm, pub_input, k, alpha, beta, gamma, keep):
        """pointers may be ints (device) or numpy arrays (host; kept alive in `keep`)"""
        def ptr(x):
            if isinstance(x, int):
                return x
            a = np.ascontiguousarray(x, dtype=np.uint64)
            keep.append(a)
            return a.ctypes.data
        q = QuotientArgs()
        for j in range(13):
            q.selectors[j] = ptr(selectors[j])
        for j in range(5):
            q.sigmas[j] = ptr(sigmas[j])
            q.wires[j] = ptr(wires[j])
        q.perm, q.pub_input = ptr(perm), ptr(pub_input)
        q.k, q.alpha, q.beta, q.gamma = ptr(np.asarray(k)), ptr(np.asarray(alpha)), ptr(np.asarray(beta)), ptr(np.asarray(gamma))
        return q

    def quotient_evals(self, selectors, sigmas, wires, perm, pub_input, k, alpha, beta, gamma) -> np.ndarray:
        """round 3 (dispatcher2.rs:434-504) on host arrays: selectors [13][m,4], sigmas / wires [5][m,4], ..."""
        keep = []
        q = self._quotient_args(selectors, sigmas, wires, perm, pub_input, k, alpha, beta, gamma, keep)
        out = np.empty((np.asarray(perm).shape[0], 4), dtype=np.uint64)
        self._ck(self.lib.dp_quotient_evals(self.h, C.byref(q), _addr(out)))
        return out

    def quotient_evals_dev(self, selectors, sigmas, wires, perm, pub_input, k, alpha, beta, gamma, out_ptr: int):
        keep = []
        q = self._quotient_args(selectors, sigmas, wires, perm, pub_input, k, alpha, beta, gamma, keep)
        self._ck(self.lib.dp_quotient_evals_dev(self.h, C.byref(q), out_ptr))

    def poly_eval(self, coeffs, point: np.ndarray, n: int | None = None) -> np.ndarray:
        """round 4: p(point); coeffs = [n,4] host array, or a device pointer with n given"""
        pt, out = np.ascontiguousarray(point, dtype=np.uint64), np.empty(4, dtype=np.uint64)
        if isinstance(coeffs, int):
            self._ck(self.lib.dp_poly_eval_dev(self.h, coeffs, n, _addr(pt), _addr(out)))
        else:
            c = np.ascontiguousarray(coeffs, dtype=np.uint64)
            self._ck(self.lib.dp_poly_eval(self.h, _addr(c) if c.size else None, c.shape[0], _addr(pt), _addr(out)))
        return out

    def poly_div_linear(self, coeffs, point: np.ndarray, n: int | None = None, out_ptr: int | None = None):
        """round 5: (quotient of p / (X - point), remainder p(point)); device form writes the quotient to out_ptr"""
        pt, rem = np.ascontiguousarray(point, dtype=np.uint64), np.empty(4, dtype=np.uint64)
        if isinstance(coeffs, int):
            self._ck(self.lib.dp_poly_div_linear_dev(self.h, coeffs, n, _addr(pt), out_ptr, _addr(rem)))
            return None, rem
        c = np.ascontiguousarray(coeffs, dtype=np.uint64)
        out = np.empty((max(c.shape[0] - 1, 0), 4), dtype=np.uint64)
        self._ck(self.lib.dp_poly_div_linear(self.h, _addr(c) if c.size else None, c.shape[0], _addr(pt), _addr(out) if out.size else None, _addr(rem)))
        return out, rem

    def poly_lincomb(self, polys, coeffs: np.ndarray, out_len: int | None = None, lens=None, out_ptr: int | None = None):
        """round 5: sum_i coeffs[i] * polys[i]; polys = host arrays, or device pointers with lens and out_ptr given"""
        cf = np.ascontiguousarray(coeffs, dtype=np.uint64)
        k = len(polys)
        if out_ptr is not None:
            ptrs, ln = (C.c_void_p * k)(*polys), (C.c_size_t * k)(*lens)
            self._ck(self.lib.dp_poly_lincomb_dev(self.h, ptrs, ln, _addr(cf), k, out_ptr, out_len))
            return None
        ps = [np.ascontiguousarray(x, dtype=np.uint64) for x in polys]
        ln = [x.shape[0] for x in ps]
        n_out = max(ln) if out_len is None else out_len
        out = np.empty((n_out, 4), dtype=np.uint64)
        ptrs = (C.c_void_p * k)(*[x.ctypes.data if x.size else None for x in ps])
        self._ck(self.lib.dp_poly_lincomb(self.h, ptrs, (C.c_size_t * k)(*ln), _addr(cf), k, _addr(out) if n_out else None, n_out))
        return out

    def get_wire(self) -> np.ndarray:
        n = C.c_size_t()
        self._ck(self.lib.dp_get_wire(self.h, None, 0, C.byref(n)))
        out = np.empty((n.value, 4), dtype=np.uint64)
        self._ck(self.lib.dp_get_wire(self.h, _addr(out), out.nbytes, C.byref(n)))
        return out

    # ---- device-pointer variants (bench)
    def msm_dev(self, start, end, scalars_ptr: int, n: int, out_ptr: int):
        self._ck(self.lib.dp_msm_dev(self.h, start, end, scalars_ptr, n, out_ptr))

    def msm_batch(self, jobs):
        """jobs: list of (start, end, scalars ndarray | host address, n_scalars); returns [144-byte arrays]"""
        k = len(jobs)
        outs = [np.zeros(G1_PROJECTIVE_BYTES, dtype=np.uint8) for _ in range(k)]
        st = (C.c_uint64 * k)(*[j[0] for j in jobs])
        en = (C.c_uint64 * k)(*[j[1] for j in jobs])
        sc = (C.c_void_p * k)(*[_addr(j[2]) for j in jobs])
        ns = (C.c_size_t * k)(*[j[3] for j in jobs])
        ou = (C.c_void_p * k)(*[_addr(o) for o in outs])
        self._ck(self.lib.dp_msm_batch(self.h, k, st, en, sc, ns, ou))
        return outs

    def msm_dev_batch(self, jobs):
        """jobs: list of (start, end, scalars_ptr, n_scalars, out_ptr)"""
        k = len(jobs)
        st = (C.c_uint64 * k)(*[j[0] for j in jobs])
        en = (C.c_uint64 * k)(*[j[1] for j in jobs])
        sc = (C.c_void_p * k)(*[j[2] for j in jobs])
        ns = (C.c_size_t * k)(*[j[3] for j in jobs])
        ou = (C.c_void_p * k)(*[j[4] for j in jobs])
        self._ck(self.lib.dp_msm_dev_batch(self.h, k, st, en, sc, ns, ou))

    def ntt_dev(self, data_ptr: int, log_n: int, is_inv: bool, is_coset: bool):
        self._ck(self.lib.dp_ntt_dev(self.h, data_ptr, log_n, int(is_inv), int(is_coset)))

    def ntt_dev_padded(self, data_ptr: int, n_valid: int, log_n: int, is_inv: bool, is_coset: bool, wait: bool = True):
        """in place on 2^log_n Fr at data_ptr whose entries from n_valid on are zero"""
        self._ck(self.lib.dp_ntt_dev_padded(self.h, data_ptr, n_valid, log_n, int(is_inv), int(is_coset), int(wait)))

    def fft_dev(self, rows_ptr: int, cols_ptr: int, is_quot: bool, is_inv: bool, is_coset: bool):
        self._ck(self.lib.dp_fft_dev(self.h, rows_ptr, cols_ptr, int(is_quot), int(is_inv), int(is_coset)))

    def fft_dev_hint_valid_cols(self, is_quot: bool, valid_cols: int):
        self._ck(self.lib.dp_fft_dev_hint_valid_cols(self.h, int(is_quot), valid_cols))

    def fft_dev_rows(self, rows_ptr: int, is_quot: bool, is_inv: bool, is_coset: bool):
        s, r, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._ck(self.lib.dp_fft_dev_rows(self.h, rows_ptr, int(is_quot), int(is_inv), int(is_coset),
                                          C.byref(s), C.byref(r), C.byref(n)))
        return s.value, r.value, n.value

    def peer_arena_create(self, arena_bytes: int) -> bytes:
        h = C.create_string_buffer(64)
        self._ck(self.lib.dp_peer_arena_create(self.h, arena_bytes, h))
        return h.raw

    def peer_attach(self, peer: int, handle: bytes):
        self._ck(self.lib.dp_peer_attach(self.h, peer, C.create_string_buffer(handle, 64)))

    def peer_ready(self) -> bool:
        return bool(self.lib.dp_peer_ready(self.h))

    def fft_dev_rows_p2p(self, rows_ptr: int, is_quot: bool, is_inv: bool, is_coset: bool):
        self._ck(self.lib.dp_fft_dev_rows_p2p(self.h, rows_ptr, int(is_quot), int(is_inv), int(is_coset)))

    def fft_dev_p2p(self, rows_ptr: int, cols_ptr: int, is_quot: bool, is_inv: bool, is_coset: bool):
        self._ck(self.lib.dp_fft_dev_p2p(self.h, rows_ptr, cols_ptr, int(is_quot), int(is_inv), int(is_coset)))

    def fft_dev_p2p_async(self, rows_ptr: int, cols_ptr: int, is_quot: bool, is_inv: bool, is_coset: bool):
        self._ck(self.lib.dp_fft_dev_p2p_async(self.h, rows_ptr, cols_ptr, int(is_quot), int(is_inv), int(is_coset)))

    def fft_dev_cols(self, cols_ptr: int):
        self._ck(self.lib.dp_fft_dev_cols(self.h, cols_ptr))

    def debug_set_limits(self, max_contig_log_k=11, max_strided_log_k=9, msm_window_bits=0):
        self._ck(self.lib.dp_debug_set_limits(self.h, max_contig_log_k, max_strided_log_k, msm_window_bits))

    def debug_set_three_pass(self, min_log_n: int):
        self._ck(self.lib.dp_debug_set_three_pass(self.h, min_log_n))

    def msm_breakdown(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self.lib.dp_last_msm_breakdown(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def msm_tuning(self):
        """dp_init's choice between the plain MSM pipeline and batched-affine tree levels:
        {"plain_ms", "affine_ms", "levels", "equal"} (equal: 1 same result, 0 different, -1 not run)"""
        a, b, lv, eq = C.c_float(), C.c_float(), C.c_int(), C.c_int()
        self._ck(self.lib.dp_msm_tuning(self.h, C.byref(a), C.byref(b), C.byref(lv), C.byref(eq)))
        out = {"plain_ms": a.value, "affine_ms": b.value, "levels": lv.value, "equal": eq.value}
        allms = (C.c_float * 4)()
        self._ck(self.lib.dp_msm_tuning_all(self.h, allms))
        if any(v > 0 for v in allms):
            out["ms_by_levels"] = [round(float(v), 4) for v in allms]
        return out

    def gen_bases(self, seed: int, n: int) -> np.ndarray:
        out = np.zeros((n, G1_AFFINE_BYTES), dtype=np.uint8)
        self._ck(self.lib.dp_debug_gen_bases(self.h, seed, n, _addr(out) if n else None))
        return out

    def gen_bases_into(self, seed: int, n: int, out_ptr: int):
        """the same, written to `out_ptr` (n * 104 B of host or device memory)"""
        self._ck(self.lib.dp_debug_gen_bases(self.h, seed, n, out_ptr))

    def init_ptr(self, bases_ptr: int, n_bases: int, domain_size: int, quot_domain_size: int):
        """PlonkSlave.init with the raw GroupAffine array at `bases_ptr` (host or device memory)"""
        self._ck(self.lib.dp_init(self.h, bases_ptr if n_bases else None, n_bases, domain_size, quot_domain_size))

    def sync(self):
        self._ck(self.lib.dp_sync(self.h))

    def last_timing(self):
        ms, n = C.c_float(), C.c_uint64()
        self.lib.dp_last_timing(self.h, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def launch_count(self) -> int:
        return int(self.lib.dp_launch_count(self.h))
