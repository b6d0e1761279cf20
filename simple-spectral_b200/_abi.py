"""ctypes mirror of include/ssb200.h (POD structs of the C ABI). No compute lives here."""
import ctypes as C

SSB_OK, SSB_ERR_DATA, SSB_ERR_ARG, SSB_ERR_UNSUPPORTED = 0, -1, -2, -3
SSB_FILTER_LINEAR, SSB_FILTER_NEAREST = 0, 1
SSB_MATERIAL_LAMBERT, SSB_MATERIAL_MIRROR = 0, 1
SSB_ALBEDO_CONSTANT, SSB_ALBEDO_TEXTURE = 0, 1
SSB_UPSAMPLE_OURS, SSB_UPSAMPLE_MENG, SSB_UPSAMPLE_JH = 1, 2, 3
SSB_RENDER_SPECTRAL, SSB_RENDER_RGB = 0, 1
SSB_SCAN_FILTERED, SSB_SCAN_LIST = 0, 1


class ssb_vertex(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("st", C.c_float * 2)]


class ssb_tri(C.Structure):
    _fields_ = [("v", ssb_vertex * 3), ("normal", C.c_float * 3)]


class ssb_quad(C.Structure):
    _fields_ = [("tri", ssb_tri * 2), ("material", C.c_uint32), ("is_light", C.c_uint32)]


class ssb_spectrum(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("n", C.c_uint32), ("low", C.c_float), ("high", C.c_float),
                ("filter", C.c_uint32)]


class ssb_material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("albedo_mode", C.c_uint32), ("albedo", ssb_spectrum),
                ("texture", C.c_uint32), ("emission", ssb_spectrum),
                ("albedo_rgb", C.c_float * 3), ("emission_rgb", C.c_float * 3)]


class ssb_texture(C.Structure):
    _fields_ = [("rgb8", C.POINTER(C.c_uint8)), ("width", C.c_uint32), ("height", C.c_uint32)]


class ssb_camera(C.Structure):
    _fields_ = [("pv_inv", C.c_double * 16), ("pos", C.c_float * 3), ("dir", C.c_float * 3)]


class ssb_scene(C.Structure):
    _fields_ = [("camera", ssb_camera), ("quads", C.POINTER(ssb_quad)), ("nquads", C.c_uint32),
                ("materials", C.POINTER(ssb_material)), ("nmaterials", C.c_uint32),
                ("textures", C.POINTER(ssb_texture)), ("ntextures", C.c_uint32)]


class ssb_meng_tables(C.Structure):
    _fields_ = [("grid", C.POINTER(C.c_int32)), ("grid_w", C.c_uint32), ("grid_h", C.c_uint32),
                ("points", C.POINTER(C.c_float)), ("npoints", C.c_uint32), ("nsamples", C.c_uint32),
                ("xy_to_uv", C.c_float * 6), ("sample_min", C.c_float), ("sample_max", C.c_float)]


class ssb_color(C.Structure):
    _fields_ = [("xbar", ssb_spectrum), ("ybar", ssb_spectrum), ("zbar", ssb_spectrum),
                ("basis_r", ssb_spectrum), ("basis_g", ssb_spectrum), ("basis_b", ssb_spectrum),
                ("xyz_to_lrgb", C.c_float * 9), ("d65_rad_Y", C.c_float),
                ("jh_scale", C.POINTER(C.c_float)), ("jh_data", C.POINTER(C.c_float)), ("jh_res", C.c_uint32),
                ("meng", C.POINTER(ssb_meng_tables))]


class ssb_options(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32),
                ("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32),
                ("sample_begin", C.c_uint32), ("sample_end", C.c_uint32),
                ("indirect_only", C.c_uint32), ("upsampling", C.c_uint32),
                ("lambda_min", C.c_float), ("lambda_max", C.c_float),
                ("max_depth", C.c_uint32), ("explicit_light_sampling", C.c_uint32),
                ("flat_field_correction", C.c_uint32), ("eps", C.c_float), ("seed", C.c_uint64),
                ("render_mode", C.c_uint32), ("n_wavelengths", C.c_uint32),
                ("keep_accumulator", C.c_uint32), ("prebaked_textures", C.c_uint32),
                ("band_height", C.c_uint32), ("band_count", C.c_uint32), ("band_index", C.c_uint32),
                ("scan_mode", C.c_uint32)]


class ssb_stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("device_ms", C.c_double), ("trace_ms", C.c_double),
                ("launches", C.c_uint32), ("reserved", C.c_uint32)]


def default_options(width, height, spp, **kw):
    """Reference defaults (stdafx.hpp:44-90); mirrors ssb_default_options()."""
    o = ssb_options()
    o.width, o.height, o.spp = width, height, spp
    o.upsampling = SSB_UPSAMPLE_OURS
    o.lambda_min, o.lambda_max = 380.0, 780.0
    o.max_depth = 10
    o.explicit_light_sampling = 1
    o.flat_field_correction = 1
    o.eps = 0.001
    o.seed = 1
    for k, v in kw.items():
        setattr(o, k, v)
    return o
