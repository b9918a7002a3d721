"""ctypes binding of libislam_pvgo.so (include/islam_pvgo.h).  No torch types cross this boundary: raw device
pointers, sizes and a cudaStream_t.  There is NO fallback: if the library is missing or a call fails, raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libislam_pvgo.so')


class IslamError(RuntimeError):
    pass


class PvgoOpts(C.Structure):
    _fields_ = [('band_max', C.c_int32), ('leaf_max', C.c_int32), ('pivot_max', C.c_int32),
                ('n_parts', C.c_int32), ('part', C.c_int32), ('reserved', C.c_int32 * 3)]


class PvgoDims(C.Structure):
    _fields_ = [('N', C.c_int32), ('E', C.c_int32), ('M', C.c_int32), ('P', C.c_int32), ('F', C.c_int32),
                ('levels', C.c_int32), ('band', C.c_int32), ('root_pivots', C.c_int32), ('max_rows', C.c_int32),
                ('max_cols', C.c_int32), ('n_shared_fronts', C.c_int32), ('bs_launches', C.c_int32),
                ('L_doubles', C.c_int64), ('U_doubles', C.c_int64), ('shared_doubles', C.c_int64),
                ('factor_flops', C.c_double)]


class LMState(C.Structure):
    _fields_ = [('loss', C.c_double), ('last', C.c_double), ('loss_trial', C.c_double), ('damping', C.c_double),
                ('radius', C.c_double), ('down', C.c_double), ('diag_scale', C.c_double), ('quality', C.c_double),
                ('denom', C.c_double), ('lin_loss', C.c_double),
                ('reject_count', C.c_int32), ('steps_done', C.c_int32), ('tries_total', C.c_int32),
                ('accepted_last', C.c_int32), ('need_linearize', C.c_int32), ('continual', C.c_int32),
                ('patience_count', C.c_int32), ('info', C.c_int32), ('cur', C.c_int32), ('active', C.c_int32),
                ('do_lin', C.c_int32), ('chol_fail', C.c_int32), ('loss_valid', C.c_int32), ('pad0', C.c_int32),
                ('pad1', C.c_int32), ('pad2', C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith('pad')}


class LMParams(C.Structure):
    _fields_ = [('radius', C.c_double), ('lm_min', C.c_double), ('lm_max', C.c_double), ('high', C.c_double),
                ('low', C.c_double), ('up', C.c_double), ('down', C.c_double), ('factor', C.c_double),
                ('tr_min', C.c_double), ('tr_max', C.c_double), ('reject', C.c_int32), ('max_steps', C.c_int32),
                ('patience', C.c_int32), ('use_scheduler', C.c_int32), ('decreasing', C.c_double)]


_P = C.c_void_p
_SIGS = {
    # name: (restype, argtypes)
    'islam_pvgo_create': (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32, _P, C.POINTER(PvgoOpts)]),
    'islam_pvgo_destroy': (None, [_P]),
    'islam_pvgo_get_dims': (C.c_int, [_P, C.POINTER(PvgoDims)]),
    'islam_lm_default_params': (None, [C.POINTER(LMParams)]),
    'islam_pvgo_set_problem': (C.c_int, [_P, _P, _P, _P, _P, _P, C.POINTER(C.c_double * 4), _P]),
    'islam_pvgo_set_state': (C.c_int, [_P, _P, _P, _P]),
    'islam_pvgo_set_reproj': (C.c_int, [_P, _P, _P, C.c_int32, C.POINTER(C.c_float * 4), C.POINTER(C.c_float * 7), C.c_double, _P]),
    'islam_pvgo_get_reproj_residuals': (C.c_int, [_P, _P, _P]),
    'islam_pvgo_get_state': (C.c_int, [_P, _P, _P, _P]),
    'islam_pvgo_linearize': (C.c_int, [_P, _P]),
    'islam_pvgo_get_residuals': (C.c_int, [_P, _P, _P, _P, _P, _P]),
    'islam_pvgo_get_normal_eq': (C.c_int, [_P, _P, _P, _P, _P, _P]),
    'islam_pvgo_solve': (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _P, C.POINTER(C.c_int32), _P]),
    'islam_pvgo_lm_reset': (C.c_int, [_P, C.POINTER(LMParams), _P]),
    'islam_pvgo_lm_try': (C.c_int, [_P, _P]),
    'islam_pvgo_lm_step': (C.c_int, [_P, C.POINTER(LMState), _P]),
    'islam_pvgo_lm_run': (C.c_int, [_P, C.POINTER(LMState), _P]),
    'islam_pvgo_get_lm_state': (C.c_int, [_P, C.POINTER(LMState), _P]),
    'islam_pvgo_profile_try': (C.c_int, [_P, C.POINTER(C.c_float * 5), _P]),
    'islam_pvgo_lm_try_begin': (C.c_int, [_P, _P]),
    'islam_pvgo_shared_buffer': (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    'islam_pvgo_lm_try_mid': (C.c_int, [_P, _P]),
    'islam_pvgo_lm_try_mid2': (C.c_int, [_P, _P]),
    'islam_pvgo_root_buffers': (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(_P),
                                          C.POINTER(C.c_int32)]),
    'islam_pvgo_root_owner': (C.c_int, [_P, C.c_int64]),
    'islam_pvgo_root_panel': (C.c_int, [_P, C.c_int64, _P]),
    'islam_pvgo_root_update': (C.c_int, [_P, C.c_int64, _P]),
    'islam_pvgo_root_update_part': (C.c_int, [_P, C.c_int64, C.c_int32, _P]),
    'islam_pvgo_root_zero_foreign': (C.c_int, [_P, _P]),
    'islam_pvgo_sums_buffer': (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    'islam_pvgo_lm_try_end': (C.c_int, [_P, _P]),
    'islam_pvgo_var_parts': (C.c_int, [_P, _P]),
    'islam_pvgo_mailbox_export': (C.c_int, [_P, _P]),
    'islam_pvgo_mailbox_connect': (C.c_int, [_P, _P]),
    'islam_pvgo_small_supported': (C.c_int, [C.c_int32, C.c_int32]),
    'islam_pvgo_small_run': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_double * 4),
                                       C.POINTER(LMParams), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'islam_pvgo_vo_loss': (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    'islam_pvgo_imu_loss': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    'islam_pvgo_align': (C.c_int, [_P, _P, _P, _P, _P]),
    'islam_imu_preintegrate': (C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, _P, C.c_float, C.c_int32, _P, _P, _P,
                                         _P, _P]),
    'islam_imu_workspace_bytes': (C.c_int64, [C.c_int32, C.c_int32]),
    'islam_scale_from_disp_flow': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    'islam_scale_workspace_bytes': (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    'islam_lie_exp': (C.c_int, [C.c_int32, _P, _P, C.c_int64, _P]),
    'islam_lie_log': (C.c_int, [C.c_int32, _P, _P, C.c_int64, _P]),
    'islam_lie_inv': (C.c_int, [C.c_int32, _P, _P, C.c_int64, _P]),
    'islam_lie_mul': (C.c_int, [C.c_int32, _P, _P, _P, C.c_int64, _P]),
    'islam_lie_act': (C.c_int, [C.c_int32, _P, _P, _P, C.c_int64, _P]),
    'islam_lie_exp_bwd': (C.c_int, [C.c_int32, _P, _P, _P, C.c_int64, _P]),
    'islam_lie_log_bwd': (C.c_int, [C.c_int32, _P, _P, _P, C.c_int64, _P]),
    'islam_lie_inv_bwd': (C.c_int, [C.c_int32, _P, _P, _P, C.c_int64, _P]),
    'islam_lie_mul_bwd': (C.c_int, [C.c_int32, _P, _P, _P, _P, C.c_int64, _P]),
    'islam_lie_act_bwd': (C.c_int, [C.c_int32, _P, _P, _P, _P, _P, C.c_int64, _P]),
    'islam_lie_cumprod': (C.c_int, [C.c_int32, _P, _P, C.c_int64, C.c_int32, _P]),
    'islam_plan_build': (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32, _P, C.POINTER(PvgoOpts)]),
    'islam_plan_free': (None, [_P]),
    'islam_plan_array': (C.c_int64, [_P, C.c_char_p, C.POINTER(_P)]),
    'islam_version': (C.c_char_p, []),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib():
    """The loaded library.  Raises IslamError if it has not been built — there is no CPU / PyTorch fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IslamError(f'{LIB_PATH} is missing: build it with `python -m islam_b200.build` '
                             '(the CUDA extension is required; there is no fallback path)')
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        kind = 'cudaError' if rc > 0 else 'invalid argument / unsupported'
        raise IslamError(f'{what} failed with code {rc} ({kind})')
