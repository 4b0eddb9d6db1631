"""ctypes binding of libinvertavatar_b200.so (the C-ABI in include/invertavatar_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised (the reference raises RuntimeError through TORCH_CHECK, torch_utils/ops/bias_act.cpp:39-55)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libinvertavatar_b200.so')

c_f32p = C.c_void_p
c_u16p = C.c_void_p
c_i32p = C.c_void_p


class StyleLayer(C.Structure):
    _fields_ = [('affine_w', c_f32p), ('affine_b', c_f32p), ('wsq', c_f32p), ('styles', c_f32p), ('dcoef', c_f32p),
                ('Cin', C.c_int32), ('Cout', C.c_int32), ('w_index', C.c_int32), ('w_dim', C.c_int32),
                ('affine_gain', C.c_float), ('style_gain', C.c_float)]


class Upfirdn2dParams(C.Structure):
    _fields_ = [('x', c_f32p), ('f', c_f32p), ('y', c_f32p),
                ('N', C.c_int32), ('C', C.c_int32), ('inH', C.c_int32), ('inW', C.c_int32),
                ('outH', C.c_int32), ('outW', C.c_int32), ('fh', C.c_int32), ('fw', C.c_int32),
                ('upx', C.c_int32), ('upy', C.c_int32), ('downx', C.c_int32), ('downy', C.c_int32),
                ('padx0', C.c_int32), ('pady0', C.c_int32), ('flip', C.c_int32), ('gain', C.c_float),
                ('xs_n', C.c_int64), ('xs_c', C.c_int64), ('xs_h', C.c_int64), ('xs_w', C.c_int64),
                ('ys_n', C.c_int64), ('ys_c', C.c_int64), ('ys_h', C.c_int64), ('ys_w', C.c_int64), ('dtype', C.c_int32)]


class FilteredLreluParams(C.Structure):
    _fields_ = [('x', C.c_void_p), ('y', C.c_void_p), ('b', C.c_void_p), ('fu', c_f32p), ('fd', c_f32p),
                ('N', C.c_int32), ('C', C.c_int32), ('inH', C.c_int32), ('inW', C.c_int32), ('outH', C.c_int32), ('outW', C.c_int32),
                ('fuw', C.c_int32), ('fuh', C.c_int32), ('fdw', C.c_int32), ('fdh', C.c_int32),
                ('up', C.c_int32), ('down', C.c_int32), ('px0', C.c_int32), ('px1', C.c_int32), ('py0', C.c_int32), ('py1', C.c_int32),
                ('gain', C.c_float), ('slope', C.c_float), ('clamp', C.c_float), ('flip', C.c_int32),
                ('xs_n', C.c_int64), ('xs_c', C.c_int64), ('xs_h', C.c_int64), ('xs_w', C.c_int64),
                ('ys_n', C.c_int64), ('ys_c', C.c_int64), ('ys_h', C.c_int64), ('ys_w', C.c_int64),
                ('dtype', C.c_int32), ('workspace', C.c_void_p), ('workspace_bytes', C.c_int64),
                ('si', C.c_void_p), ('sx', C.c_int32), ('sy', C.c_int32), ('write_signs', C.c_int32), ('so', C.c_void_p)]


class ModsplitParams(C.Structure):
    _fields_ = [('x', c_f32p), ('x_ld', C.c_int64), ('styles', c_f32p), ('cond', c_f32p), ('cond_ld', C.c_int64),
                ('cond_alpha', c_f32p), ('hi', c_u16p), ('lo', c_u16p),
                ('B', C.c_int32), ('HW', C.c_int32), ('C', C.c_int32), ('C_pad', C.c_int32), ('out_img_pix', C.c_int64), ('fmt', C.c_int32)]


class Emit(C.Structure):
    _fields_ = [('out32', c_f32p), ('out32_ld', C.c_int64),
                ('hi1', c_u16p), ('lo1', c_u16p), ('s1', c_f32p), ('c1_pad', C.c_int32),
                ('hi2', c_u16p), ('lo2', c_u16p), ('s2', c_f32p), ('c2_pad', C.c_int32),
                ('rgb_out', c_f32p), ('rgb_w', c_f32p), ('rgb_s', c_f32p), ('rgb_n', C.c_int32),
                ('e1_img_pix', C.c_int64), ('fmt1', C.c_int32), ('fmt2', C.c_int32)]


class ConvParams(C.Structure):
    _fields_ = [('a_hi', c_u16p), ('a_lo', c_u16p), ('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('Cin_pad', C.c_int32),
                ('w_hi', c_u16p), ('w_lo', c_u16p), ('Cout', C.c_int32), ('Cout_pad', C.c_int32), ('n_taps_total', C.c_int32),
                ('GH', C.c_int32), ('GW', C.c_int32), ('ntaps', C.c_int32),
                ('dy', C.c_int32 * 9), ('dx', C.c_int32 * 9), ('wtap', C.c_int32 * 9),
                ('OH', C.c_int32), ('OW', C.c_int32), ('sy', C.c_int32), ('sx', C.c_int32), ('py', C.c_int32), ('px', C.c_int32),
                ('mode', C.c_int32), ('dcoef', c_f32p), ('noise', c_f32p), ('noise_strength', c_f32p), ('bias', c_f32p),
                ('noise_bstride', C.c_int64),
                ('act', C.c_int32), ('alpha', C.c_float), ('gain', C.c_float), ('clamp', C.c_float),
                ('emit', Emit),
                ('groups', C.c_int32), ('imgs_per_group', C.c_int32), ('noise_gstride', C.c_int64),
                ('img_prev', c_f32p), ('a_img_rows', C.c_int32),
                ('splitk_ws', c_f32p), ('splitk_ws_bytes', C.c_int64), ('splitk_counters', c_i32p), ('splitk_n_counters', C.c_int32),
                ('op_fmt', C.c_int32),
                ('slope', c_f32p)]


class FirParams(C.Structure):
    _fields_ = [('raw', c_f32p), ('B', C.c_int32), ('RH', C.c_int32), ('RW', C.c_int32), ('C', C.c_int32),
                ('fir', c_f32p), ('OH', C.c_int32), ('OW', C.c_int32),
                ('dcoef', c_f32p), ('noise', c_f32p), ('noise_strength', c_f32p), ('bias', c_f32p),
                ('noise_bstride', C.c_int64),
                ('act', C.c_int32), ('alpha', C.c_float), ('gain', C.c_float), ('clamp', C.c_float),
                ('emit', Emit),
                ('groups', C.c_int32), ('imgs_per_group', C.c_int32), ('noise_gstride', C.c_int64)]


class TorgbParams(C.Structure):
    _fields_ = [('raw', c_f32p), ('raw_ld', C.c_int64), ('bias', c_f32p), ('clamp', C.c_float), ('img_prev', c_f32p),
                ('img_out', c_f32p), ('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C', C.c_int32),
                ('out_nchw', C.c_int32), ('groups', C.c_int32), ('imgs_per_group', C.c_int32),
                ('peer_out', c_f32p * 8), ('n_peers', C.c_int32), ('peer_offset', C.c_int64), ('mc_out', c_f32p)]


class ResizeParams(C.Structure):
    _fields_ = [('inp', c_f32p), ('in_ld', C.c_int64), ('in_H', C.c_int32), ('in_W', C.c_int32),
                ('in_y0', C.c_int32), ('in_x0', C.c_int32),
                ('out', c_f32p), ('out_ld', C.c_int64), ('out_H', C.c_int32), ('out_W', C.c_int32),
                ('out_y0', C.c_int32), ('out_x0', C.c_int32),
                ('B', C.c_int32), ('C', C.c_int32), ('oh', C.c_int32), ('ow', C.c_int32),
                ('y_start', c_i32p), ('y_count', c_i32p), ('y_w', c_f32p), ('y_max_taps', C.c_int32),
                ('x_start', c_i32p), ('x_count', c_i32p), ('x_w', c_f32p), ('x_max_taps', C.c_int32)]


class LerpParams(C.Structure):
    _fields_ = [('a', c_f32p), ('a_ld', C.c_int64), ('a_row', C.c_int64), ('a_batch', C.c_int64),
                ('b', c_f32p), ('b_ld', C.c_int64), ('b_row', C.c_int64), ('b_batch', C.c_int64),
                ('alpha', c_f32p), ('al_ld', C.c_int64), ('al_row', C.c_int64), ('al_batch', C.c_int64),
                ('out', c_f32p), ('o_ld', C.c_int64), ('o_row', C.c_int64), ('o_batch', C.c_int64),
                ('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C', C.c_int32)]


class RenderParams(C.Structure):
    _fields_ = [('planes', c_f32p), ('plane_px_ld', C.c_int64), ('B', C.c_int32), ('PH', C.c_int32), ('PW', C.c_int32),
                ('cam', c_f32p), ('cam_ld', C.c_int64), ('rays_o', c_f32p), ('rays_d', c_f32p), ('res', C.c_int32), ('Dc', C.c_int32), ('Df', C.c_int32),
                ('jitter', c_f32p), ('u', c_f32p), ('box_warp', C.c_float), ('white_back', C.c_int32),
                ('near_far', c_f32p), ('w1', c_f32p), ('b1', c_f32p), ('w2', c_f32p), ('b2', c_f32p),
                ('feat', c_f32p), ('depth', c_f32p), ('wsum', c_f32p), ('depth_minmax', c_f32p), ('scratch', C.c_void_p), ('mlp_fmt', C.c_int32), ('planes_fmt', C.c_int32)]


class StitchParams(C.Structure):
    _fields_ = [('planes', c_f32p), ('planes_ld', C.c_int64), ('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C', C.c_int32),
                ('stitch', c_f32p), ('alpha', c_f32p), ('y0', C.c_int32), ('x0', C.c_int32), ('wh', C.c_int32), ('ww', C.c_int32),
                ('out', C.c_void_p), ('out_fmt', C.c_int32)]


class BlendshapeParams(C.Structure):
    _fields_ = [('neutral', c_f32p), ('exp_basis_t', c_f32p), ('exp', c_f32p), ('B', C.c_int32), ('NV', C.c_int32), ('exp_dims', C.c_int32),
                ('eye0', C.c_int32), ('eye1', C.c_int32), ('eye2', C.c_int32), ('eye_rot', c_f32p), ('eye_centre', c_f32p),
                ('M', C.c_float * 12), ('verts', c_f32p)]


class OrthoRasterParams(C.Structure):
    _fields_ = [('verts', c_f32p), ('B', C.c_int32), ('NV', C.c_int32), ('tri', c_i32p), ('F', C.c_int32), ('attr', c_f32p),
                ('size', C.c_int32), ('cam_z', C.c_float), ('blur_radius', C.c_float),
                ('crop_x', C.c_int32), ('crop_y', C.c_int32), ('crop_w', C.c_int32), ('crop_h', C.c_int32),
                ('zbuf', C.c_void_p), ('out', c_f32p), ('pix_to_face', c_i32p)]


class RasterLevelParams(C.Structure):
    _fields_ = [('tex', c_f32p), ('Ht', C.c_int32), ('Wt', C.c_int32), ('C', C.c_int32),
                ('uv', c_f32p), ('uv_ld', C.c_int64), ('UH', C.c_int32), ('UW', C.c_int32),
                ('tmp', c_f32p),
                ('stat', c_f32p), ('stat_ld', C.c_int64), ('SH', C.c_int32), ('SW', C.c_int32), ('sy0', C.c_int32), ('sx0', C.c_int32),
                ('alpha', c_f32p),
                ('out', c_f32p), ('out_ld', C.c_int64),
                ('B', C.c_int32), ('r', C.c_int32),
                ('ux_start', c_i32p), ('ux_count', c_i32p), ('ux_w', c_f32p), ('ux_max_taps', C.c_int32),
                ('uy_start', c_i32p), ('uy_count', c_i32p), ('uy_w', c_f32p), ('uy_max_taps', C.c_int32),
                ('sx_start', c_i32p), ('sx_count', c_i32p), ('sx_w', c_f32p), ('sx_max_taps', C.c_int32),
                ('sy_start', c_i32p), ('sy_count', c_i32p), ('sy_w', c_f32p), ('sy_max_taps', C.c_int32)]


class View(C.Structure):
    _fields_ = [('p', c_f32p), ('C', C.c_int32), ('ps', C.c_int32),
                ('s_c', C.c_int64), ('s_pix', C.c_int64), ('s_row', C.c_int64), ('s_img', C.c_int64)]


class EncPrepParams(C.Structure):
    _fields_ = [('src', View * 4), ('nsrc', C.c_int32),
                ('scale', c_f32p), ('shift', c_f32p), ('slope', c_f32p), ('lrelu', C.c_float),
                ('hi', c_u16p), ('lo', c_u16p), ('out32', c_f32p),
                ('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C_pad', C.c_int32)]


class EncAffineParams(C.Structure):
    _fields_ = [('x', View), ('scale', c_f32p), ('shift', c_f32p), ('slope1', c_f32p), ('slope2', c_f32p),
                ('act', C.c_int32), ('alpha', C.c_float),
                ('gate', c_f32p),
                ('res', View), ('res_scale', c_f32p), ('res_shift', c_f32p),
                ('y', c_f32p), ('y_ld', C.c_int64),
                ('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C', C.c_int32),
                ('e_scale', c_f32p), ('e_shift', c_f32p), ('e_hi', c_u16p), ('e_lo', c_u16p), ('e_C_pad', C.c_int32)]


class EncIm2colParams(C.Structure):
    _fields_ = [('src', View * 4), ('nsrc', C.c_int32),
                ('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('k', C.c_int32), ('stride', C.c_int32), ('pad', C.c_int32),
                ('OH', C.c_int32), ('OW', C.c_int32),
                ('hi', c_u16p), ('lo', c_u16p), ('K_pad', C.c_int32)]


class AttentionParams(C.Structure):
    _fields_ = [('q', c_f32p), ('k', c_f32p), ('v', c_f32p), ('q_ld', C.c_int64), ('k_ld', C.c_int64), ('v_ld', C.c_int64),
                ('q_bias', c_f32p), ('k_bias', c_f32p), ('v_bias', c_f32p),
                ('B', C.c_int32), ('heads', C.c_int32), ('head_dim', C.c_int32), ('Nq', C.c_int32), ('Nk', C.c_int32), ('scale', C.c_float),
                ('out32', c_f32p), ('out32_ld', C.c_int64),
                ('hi', c_u16p), ('lo', c_u16p), ('C_pad', C.c_int32)]


class AttentionTcParams(C.Structure):
    _fields_ = [('q_hi', c_u16p), ('q_lo', c_u16p), ('q_ld', C.c_int64),
                ('kv_hi', c_u16p), ('kv_lo', c_u16p), ('kv_ld', C.c_int64),
                ('B', C.c_int32), ('heads', C.c_int32), ('head_dim', C.c_int32), ('Nq', C.c_int32), ('Nk', C.c_int32), ('scale', C.c_float),
                ('out32', c_f32p), ('out32_ld', C.c_int64),
                ('hi', c_u16p), ('lo', c_u16p), ('C_pad', C.c_int32)]


# name -> (restype, argtypes); every symbol include/invertavatar_b200.h declares
SIGNATURES = {
    'ia_abi_version': (C.c_int, []),
    'ia_last_error': (C.c_char_p, []),
    'ia_set_device': (C.c_int, [C.c_int]),
    'ia_launch_count': (C.c_int64, []),
    'ia_reset_launch_count': (None, []),
    'ia_profile_begin': (C.c_int, []),
    'ia_profile_report': (C.c_int64, [C.c_char_p, C.c_int64]),
    'ia_bias_act': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    'ia_upfirdn2d': (C.c_int, [C.POINTER(Upfirdn2dParams), C.c_void_p]),
    'ia_filtered_lrelu_workspace': (C.c_int64, [C.POINTER(FilteredLreluParams)]),
    'ia_filtered_lrelu': (C.c_int, [C.POINTER(FilteredLreluParams), C.c_void_p]),
    'ia_filtered_lrelu_act': (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p]),
    'ia_fully_connected': (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                    C.c_int, C.c_float, C.c_float, C.c_int64, C.c_int64, C.c_void_p]),
    'ia_normalize_2nd_moment': (C.c_int, [c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_float, C.c_int64, C.c_int64, C.c_void_p]),
    'ia_broadcast_truncate': (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    'ia_styles': (C.c_int, [C.c_void_p, C.POINTER(StyleLayer), C.c_int32, c_f32p, C.c_int32, C.c_int32, C.c_void_p]),
    'ia_modsplit': (C.c_int, [C.POINTER(ModsplitParams), C.c_void_p]),
    'ia_pack_conv_weight': (C.c_int, [c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_u16p, c_u16p, c_f32p, C.c_int32, C.c_void_p]),
    'ia_conv_tc': (C.c_int, [C.POINTER(ConvParams), C.c_void_p]),
    'ia_conv_tc_phases': (C.c_int, [C.POINTER(ConvParams), C.c_int32, C.c_void_p]),
    'ia_conv_simt': (C.c_int, [C.POINTER(ConvParams), C.c_void_p]),
    'ia_fir_epilogue': (C.c_int, [C.POINTER(FirParams), C.c_void_p]),
    'ia_torgb_finish': (C.c_int, [C.POINTER(TorgbParams), C.c_void_p]),
    'ia_fill_mouth': (C.c_int, [c_f32p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    'ia_grid_sample': (C.c_int, [c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, c_f32p, C.c_int64, C.c_int32, C.c_int32, c_f32p, C.c_int64, C.c_void_p]),
    'ia_resize_aa': (C.c_int, [C.POINTER(ResizeParams), C.c_void_p]),
    'ia_lerp_alpha': (C.c_int, [C.POINTER(LerpParams), C.c_void_p]),
    'ia_raster_level': (C.c_int, [C.POINTER(RasterLevelParams), C.c_void_p]),
    'ia_ray_bounds': (C.c_int, [c_f32p, C.c_int64, C.c_int32, c_f32p, C.c_void_p]),
    'ia_ray_bounds_from_origins': (C.c_int, [c_f32p, C.c_int64, c_f32p, C.c_void_p]),
    'ia_render': (C.c_int, [C.POINTER(RenderParams), C.c_void_p]),
    'ia_render_scratch_bytes': (C.c_int64, []),
    'ia_stitch_planes': (C.c_int, [C.POINTER(StitchParams), C.c_void_p]),
    'ia_ray_march': (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    'ia_depth_clamp': (C.c_int, [c_f32p, C.c_int64, c_f32p, C.c_void_p]),
    'ia_ray_sampler': (C.c_int, [c_f32p, C.c_int64, C.c_int32, C.c_int32, c_f32p, c_f32p, C.c_void_p]),
    'ia_enc_chan_stats': (C.c_int, [C.POINTER(View), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    'ia_enc_bn_fold': (C.c_int, [C.c_void_p, C.c_int64, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, C.c_float, C.c_float, C.c_int32,
                                c_f32p, c_f32p, C.c_void_p]),
    'ia_enc_bn_stats_fold': (C.c_int, [C.POINTER(View), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, c_i32p, c_f32p, c_f32p, c_f32p, c_f32p,
                                      C.c_float, C.c_float, c_f32p, c_f32p, C.c_void_p]),
    'ia_enc_prep': (C.c_int, [C.POINTER(EncPrepParams), C.c_void_p]),
    'ia_enc_affine_act': (C.c_int, [C.POINTER(EncAffineParams), C.c_void_p]),
    'ia_enc_global_pool': (C.c_int, [C.POINTER(View), c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, c_f32p, C.c_void_p]),
    'ia_enc_se_gate': (C.c_int, [C.POINTER(View), c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, c_f32p, c_f32p, C.c_int32, c_f32p, c_i32p, c_f32p,
                                C.c_void_p]),
    'ia_enc_avgpool': (C.c_int, [C.POINTER(View), C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_f32p, C.c_void_p]),
    'ia_enc_upsample_add': (C.c_int, [c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_f32p, C.c_int32, C.c_int32, c_f32p, C.c_void_p]),
    'ia_enc_gru_gate': (C.c_int, [C.c_int32, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_int32, C.c_void_p]),
    'ia_mesh_coeffs': (C.c_int, [c_f32p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    'ia_mesh_eye_centres': (C.c_int, [c_f32p, C.c_int32, C.c_int32, C.c_int32, c_f32p, C.c_void_p]),
    'ia_blendshape': (C.c_int, [C.POINTER(BlendshapeParams), C.c_void_p]),
    'ia_ortho_raster_scratch_bytes': (C.c_int64, [C.c_int32, C.c_int32]),
    'ia_ortho_raster': (C.c_int, [C.POINTER(OrthoRasterParams), C.c_void_p]),
    'ia_layout_grid_u8': (C.c_int, [c_f32p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_void_p]),
    'ia_sft_half': (C.c_int, [c_f32p, C.c_int64, C.POINTER(View), C.POINTER(View), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    'ia_enc_im2col': (C.c_int, [C.POINTER(EncIm2colParams), C.c_void_p]),
    'ia_layer_norm': (C.c_int, [c_f32p, C.c_int64, c_f32p, c_f32p, c_f32p, C.c_float, C.c_int64, C.c_int32, c_f32p, C.c_int64, c_u16p, c_u16p,
                               C.c_int32, C.c_void_p]),
    'ia_attention': (C.c_int, [C.POINTER(AttentionParams), C.c_void_p]),
    'ia_attention_tc': (C.c_int, [C.POINTER(AttentionTcParams), C.c_void_p]),
    'ia_dwconv_gelu': (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_f32p, c_u16p, c_u16p, C.c_int32,
                                C.c_void_p]),
}

ABI_VERSION = 3
_lib = None


def lib():
    """Load (once) and return the shared library; fail loudly when it is missing or stale."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -m invertavatar_b200.build` '
                '(there is no CPU or PyTorch fallback for the CUDA hot path)')
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        if l.ia_abi_version() != ABI_VERSION:
            raise RuntimeError('libinvertavatar_b200.so ABI version mismatch; rebuild it')
        _lib = l
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().ia_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'{what or "invertavatar_b200"} failed (rc={rc}): {msg}')
