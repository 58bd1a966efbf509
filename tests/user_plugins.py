"""User-written plugins: OpenCL-C fragments against the reference's kernel API.

These classes follow the reference's plugin protocol (``cl_type`` /
``cl_declaration`` / ``cl_implementation`` / ``cl_options`` / ``cl_pack``,
xopto/mcbase/mcobject.py:29-170) and carry *no* ``cu_type``: what a user of
PyXOpto who wrote their own phase function, source or detector has.  Each factory
takes the simulator module (``xopto.mcml.mc`` or ``pyxopto_b200.mcml.mc``) and
derives the class from that package's base classes, so the very same fragment
text runs through the reference's OpenCL-C kernel (golden vectors, this container
only) and through ``csrc/kernels/xo_clcompat*.cuh`` here.

The fragment text is written for this test-suite (it is not reference code).
``UserHg`` / ``UserPencil`` / ``UserRadial`` restate the arithmetic of Hg, Line at
normal incidence and Radial, so the results must equal those of the built-ins
bit for bit; ``UserCubic`` is a phase function the reference does not ship.
"""
import functools

import numpy as np


def _cltypes(mc):
    if mc.__name__.startswith('xopto'):
        from xopto.mcbase import cltypes
    else:
        from pyxopto_b200.cl import cltypes
    return cltypes


@functools.lru_cache(maxsize=None)
def _user_hg_class(mc):
    cltypes = _cltypes(mc)

    class UserHg(mc.mcpf.PfBase):
        """Henyey-Greenstein, user-written."""
        @staticmethod
        def cl_type(mc_):
            class ClUserHg(cltypes.Structure):
                _fields_ = [('g', mc_.types.mc_fp_t)]
            return ClUserHg

        @staticmethod
        def cl_declaration(mc_):
            return 'struct MC_STRUCT_ATTRIBUTES McPf{ mc_fp_t g; };\n' \
                   'void dbg_print_pf(const McPf *pf);\n'

        @staticmethod
        def cl_implementation(mc_):
            return '''
void dbg_print_pf(const McPf *pf) {
	dbg_print("user-written Hg:");
	dbg_print_float(INDENT "g:", pf->g);
};

inline mc_fp_t mcsim_pf_sample_angles(McSim *mcsim, mc_fp_t *azimuth){
	__mc_pf_mem const McPf *pf = mcsim_current_pf(mcsim);
	mc_fp_t g = pf->g;
	mc_fp_t xi, ratio, cos_theta;

	*azimuth = FP_2PI*mcsim_random(mcsim);
	xi = mcsim_random(mcsim);
	ratio = mc_fdiv(FP_1 - g*g, FP_1 + g*(FP_2*xi - FP_1));
	cos_theta = mc_fdiv(FP_1 + g*g - ratio*ratio, FP_2*g);
	if (g == FP_0)
		cos_theta = FP_1 - FP_2*mcsim_random(mcsim);

	return mc_fmax(mc_fmin(cos_theta, FP_1), -FP_1);
};
'''

        def __init__(self, g):
            super().__init__()
            self.g = float(g)

        def cl_pack(self, mc_, target=None):
            if target is None:
                target = self.cl_type(mc_)()
            target.g = self.g
            return target

        def todict(self):
            return {'type': 'UserHg', 'g': self.g}

    return UserHg


def user_hg(mc, g):
    return _user_hg_class(mc)(g)


@functools.lru_cache(maxsize=None)
def _user_cubic_class(mc):
    """p(cos t) ~ 1 + s cos^2 t (s = 1: Rayleigh), sampled by solving the cubic
    CDF with Cardano's formula - exercises mc_cbrt / mc_sqrt / mc_fclip."""
    cltypes = _cltypes(mc)

    class UserCubic(mc.mcpf.PfBase):
        @staticmethod
        def cl_type(mc_):
            class ClUserCubic(cltypes.Structure):
                _fields_ = [('s', mc_.types.mc_fp_t), ('inv_s', mc_.types.mc_fp_t)]
            return ClUserCubic

        @staticmethod
        def cl_declaration(mc_):
            return 'struct MC_STRUCT_ATTRIBUTES McPf{ mc_fp_t s; mc_fp_t inv_s; };\n' \
                   'void dbg_print_pf(const McPf *pf);\n'

        @staticmethod
        def cl_implementation(mc_):
            # CDF: (mu + s mu^3/3 + 1 + s/3) / (2 + 2s/3) = xi
            #  ->  mu^3 + (3/s) mu - (3/s)(2 xi - 1)(1 + s/3) = 0   (depressed cubic)
            return '''
void dbg_print_pf(const McPf *pf) {
	dbg_print("user-written cubic pf:");
	dbg_print_float(INDENT "s:", pf->s);
};

inline mc_fp_t mcsim_pf_sample_angles(McSim *mcsim, mc_fp_t *azimuth){
	mc_fp_t s = mcsim_current_pf(mcsim)->s;
	mc_fp_t inv_s = mcsim_current_pf(mcsim)->inv_s;
	mc_fp_t xi, p, q, d, mu;

	*azimuth = FP_2PI*mcsim_random(mcsim);
	xi = mcsim_random(mcsim);
	if (s == FP_0)
		return FP_1 - FP_2*xi;
	p = inv_s;                                          /* (3/s)/3 */
	q = FP_0p5*FP_LITERAL(3.0)*inv_s*(FP_2*xi - FP_1)*(FP_1 + s*FP_LITERAL(0.3333333333));
	d = mc_sqrt(q*q + p*p*p);
	mu = mc_cbrt(q + d) + mc_cbrt(q - d);
	return mc_fclip(mu, -FP_1, FP_1);
};
'''

        def __init__(self, strength):
            super().__init__()
            self.s = float(strength)

        def cl_pack(self, mc_, target=None):
            if target is None:
                target = self.cl_type(mc_)()
            target.s = self.s
            target.inv_s = 1.0/self.s if self.s != 0.0 else 0.0
            return target

        def todict(self):
            return {'type': 'UserCubic', 'strength': self.s}

    return UserCubic


def user_cubic(mc, strength=1.0):
    return _user_cubic_class(mc)(strength)


@functools.lru_cache(maxsize=None)
def _user_pencil_class(mc):
    """Pencil beam at normal incidence; specular reflectance packed by the host."""
    cltypes = _cltypes(mc)

    class UserPencil(mc.mcsource.Source):
        @staticmethod
        def cl_type(mc_):
            T = mc_.types
            class ClUserPencil(cltypes.Structure):
                _fields_ = [('position', T.mc_point3f_t), ('reflectance', T.mc_fp_t)]
            return ClUserPencil

        @staticmethod
        def cl_declaration(mc_):
            return 'struct MC_STRUCT_ATTRIBUTES McSource{ mc_point3f_t position; mc_fp_t reflectance; };\n'

        @staticmethod
        def cl_implementation(mc_):
            return '''
void dbg_print_source(__mc_source_mem const McSource *src){
	dbg_print("user-written pencil beam:");
	dbg_print_point3f(INDENT "position:", &src->position);
};

inline void mcsim_launch(McSim *mcsim){
	__mc_source_mem const McSource *src = mcsim_source(mcsim);
	mc_point3f_t down = {FP_0, FP_0, FP_1};

	mcsim_set_position(mcsim, &src->position);
	mcsim_set_direction(mcsim, &down);
	mcsim_set_weight(mcsim, FP_1 - src->reflectance);
	#if MC_USE_SPECULAR_DETECTOR
	{
		mc_point3f_t up = {FP_0, FP_0, -FP_1};
		mcsim_specular_detector_deposit(
			mcsim, mcsim_position(mcsim), &up, src->reflectance);
	}
	#endif
	mcsim_set_current_layer_index(mcsim, 1);
};
'''

        def __init__(self, x, y):
            super().__init__()
            self.position = np.array([x, y, 0.0])

        def cl_pack(self, mc_, target=None):
            if target is None:
                target = self.cl_type(mc_)()
            n1, n2 = float(mc_.layer(0).n), float(mc_.layer(1).n)
            target.position.fromarray(self.position)
            target.reflectance = ((n1 - n2)/(n1 + n2))**2
            return target, None, None

        def todict(self):
            return {'type': 'UserPencil', 'x': float(self.position[0]),
                    'y': float(self.position[1])}

    return UserPencil


def user_pencil(mc, x=0.0, y=0.0):
    return _user_pencil_class(mc)(x, y)


@functools.lru_cache(maxsize=None)
def _user_radial_class(mc):
    """Concentric-ring detector around the z axis (the arithmetic of Radial)."""
    cltypes = _cltypes(mc)

    class UserRadial(mc.mcdetector.Detector):
        def cl_type(self, mc_):
            T = mc_.types
            class ClUserRadial(cltypes.Structure):
                _fields_ = [('r_min', T.mc_fp_t), ('inv_dr', T.mc_fp_t),
                            ('cos_min', T.mc_fp_t), ('n', T.mc_size_t),
                            ('offset', T.mc_size_t)]
            return ClUserRadial

        def cl_declaration(self, mc_):
            Loc = self.location.capitalize()
            return 'struct MC_STRUCT_ATTRIBUTES Mc{}Detector{{ mc_fp_t r_min; mc_fp_t inv_dr; ' \
                   'mc_fp_t cos_min; mc_size_t n; mc_size_t offset; }};\n'.format(Loc)

        def cl_implementation(self, mc_):
            loc = self.location
            Loc = loc.capitalize()
            return '''
void dbg_print_{loc}_detector(__mc_detector_mem const Mc{Loc}Detector *det){{
	dbg_print("user-written ring detector:");
	dbg_print_size_t(INDENT "n:", det->n);
}};

inline void mcsim_{loc}_detector_deposit(
		McSim *mcsim, mc_point3f_t const *pos, mc_point3f_t const *dir, mc_fp_t weight){{
	__global mc_accu_t *address;
	__mc_detector_mem const struct Mc{Loc}Detector *det = mcsim_{loc}_detector(mcsim);
	mc_fp_t r = mc_sqrt(pos->x*pos->x + pos->y*pos->y);
	mc_int_t ring = mc_clip(mc_int((r - det->r_min)*det->inv_dr), 0, (mc_int_t)det->n - 1);
	uint32_t ui32w = weight_to_int(weight)*(det->cos_min <= mc_fabs(dir->z));

	address = mcsim_accumulator_buffer_ex(mcsim, det->offset + ring);
	if (ui32w > 0)
		accumulator_deposit(address, ui32w);
}};
'''.format(loc=loc, Loc=Loc)

        def __init__(self, axis, cosmin):
            super().__init__(np.zeros((axis.n,)), 0)
            self._axis = axis
            self._cosmin = float(cosmin)

        def cl_pack(self, mc_, target=None):
            if target is None:
                target = self.cl_type(mc_)()
            allocation = mc_.cl_allocate_rw_accumulator_buffer(self, self.shape)
            target.offset = allocation.offset
            target.r_min = self._axis.start
            target.inv_dr = 1.0/self._axis.step
            target.cos_min = self._cosmin
            target.n = self._axis.n
            return target

        def todict(self):
            return {'type': 'UserRadial'}

    return UserRadial


def user_radial(mc, axis, cosmin=0.0):
    return _user_radial_class(mc)(axis, cosmin)


@functools.lru_cache(maxsize=None)
def _user_depth_class(mc):
    """Deposited weight per depth slice - a fluence plugin a user wrote: 1-D, its
    own packed struct, its own result conversion."""
    cltypes = _cltypes(mc)

    class UserDepth(mc.McObject):
        @staticmethod
        def cl_type(mc_):
            T = mc_.types
            class ClUserDepth(cltypes.Structure):
                _fields_ = [('z_min', T.mc_fp_t), ('inv_dz', T.mc_fp_t),
                            ('n', T.mc_size_t), ('offset', T.mc_size_t), ('k', T.mc_int_t)]
            return ClUserDepth

        @staticmethod
        def cl_declaration(mc_):
            return 'struct MC_STRUCT_ATTRIBUTES McFluence{ mc_fp_t z_min; mc_fp_t inv_dz; ' \
                   'mc_size_t n; mc_size_t offset; mc_int_t k; };\n'

        @staticmethod
        def cl_implementation(mc_):
            return '''
void dbg_print_fluence(__mc_fluence_mem const McFluence *fluence){
	dbg_print("user-written depth profile:");
	dbg_print_size_t(INDENT "n:", fluence->n);
};

inline void mcsim_fluence_deposit_at(
		McSim *mcsim, mc_point3f_t const *position, mc_fp_t weight){
	__mc_fluence_mem McFluence const *fluence = mcsim_fluence(mcsim);
	mc_fp_t indexf = (position->z - fluence->z_min)*fluence->inv_dz;
	if (indexf >= FP_0 && indexf < fluence->n){
		mc_size_t index = mc_uint(indexf);
		uint32_t ui32w = (uint32_t)(weight*fluence->k + FP_0p5);
		mcsim_fluence_weight_deposit_ll(mcsim, fluence->offset + index, ui32w);
	};
};
'''

        def cl_options(self, mc_):
            return [('MC_USE_FLUENCE', True), ('MC_FLUENCE_MODE_RATE', False)]

        def __init__(self, zaxis, k=0x7FFFFF):
            super().__init__()
            if isinstance(zaxis, UserDepth):
                other = zaxis
                zaxis, k = other._axis, other._k
            self._axis, self._k = zaxis, int(k)
            self._raw = np.zeros((zaxis.n,))
            self._nphotons = 0

        raw = property(lambda self: self._raw)
        nphotons = property(lambda self: self._nphotons)
        k = property(lambda self: self._k)
        shape = property(lambda self: (self._axis.n,))
        mode = 'deposition'

        def cl_pack(self, mc_, target=None):
            if target is None:
                target = self.cl_type(mc_)()
            allocation = mc_.cl_allocate_rw_accumulator_buffer(self, self.shape)
            target.offset = allocation.offset
            target.z_min = self._axis.start
            target.inv_dz = 1.0/self._axis.step
            target.n = self._axis.n
            target.k = self._k
            return target

        def update_data(self, mc_, accumulators, nphotons, **kwargs):
            self._raw += np.reshape(accumulators[0], self.shape)*(1.0/self._k)
            self._nphotons += int(nphotons)

        def todict(self):
            return {'type': 'UserDepth'}

    return UserDepth


def user_depth(mc, zaxis):
    return _user_depth_class(mc)(zaxis)


# ---- sample surface layouts (mcml/mcsurface/base.py:265-287) --------------------------------
@functools.lru_cache(maxsize=None)
def _user_reflector_class(mc):
    """Diffuse / specular reflector covering a whole sample surface (the arithmetic of
    LambertianReflector), user-written."""
    cltypes = _cltypes(mc)

    class UserReflector(mc.mcsurface.SurfaceLayoutAny):
        def cl_type(self, mc_):
            T = mc_.types
            class ClUserReflector(cltypes.Structure):
                _fields_ = [('reflectance', T.mc_fp_t), ('specular', T.mc_fp_t)]
            return ClUserReflector

        def cl_declaration(self, mc_):
            return 'struct MC_STRUCT_ATTRIBUTES Mc{}SurfaceLayout{{ mc_fp_t reflectance; ' \
                   'mc_fp_t specular; }};\n'.format(self.location.capitalize())

        def cl_implementation(self, mc_):
            loc = self.location
            Loc = loc.capitalize()
            return '''
void dbg_print_{loc}_surface_layout(__mc_surface_mem const Mc{Loc}SurfaceLayout *layout){{
	dbg_print("user-written reflector:");
	dbg_print_float(INDENT "reflectance:", layout->reflectance);
}};

inline int mcsim_{loc}_surface_layout_handler(McSim *mcsim, mc_fp_t *n2, mc_fp_t *cc){{
	__mc_surface_mem const struct Mc{Loc}SurfaceLayout *layout = mcsim_{loc}_surface_layout(mcsim);
	mc_fp_t sin_fi, cos_fi, sin_theta, cos_theta;

	if (mcsim_random(mcsim) > layout->specular){{
		sin_theta = mc_sqrt(mcsim_random(mcsim));
		cos_theta = mc_sqrt(FP_1 - sin_theta*sin_theta);
		mc_sincos(mcsim_random(mcsim)*FP_2PI, &sin_fi, &cos_fi);
		mcsim_set_direction_coordinates(mcsim, cos_fi*sin_theta, sin_fi*sin_theta,
			mc_fsign(-mcsim_direction_z(mcsim))*cos_theta);
	}} else {{
		mcsim_reverse_direction_z(mcsim);
	}};
	mcsim_set_weight(mcsim, mcsim_weight(mcsim)*layout->reflectance);
	return MC_REFLECTED;
}};
'''.format(loc=loc, Loc=Loc)

        def __init__(self, reflectance, specular=0.0):
            super().__init__()
            self._reflectance, self._specular = float(reflectance), float(specular)

        def cl_pack(self, mc_, target=None):
            if target is None:
                target = self.cl_type(mc_)()
            target.reflectance = self._reflectance
            target.specular = self._specular
            return target

        def todict(self):
            return {'type': 'UserReflector', 'reflectance': self._reflectance,
                    'specular': self._specular}

    return UserReflector


def user_reflector(mc, reflectance, specular=0.0):
    return _user_reflector_class(mc)(reflectance, specular)


@functools.lru_cache(maxsize=None)
def _user_window_class(mc):
    """A layout the reference does not ship: a disc of radius ``r`` around the z axis is an
    ideally anti-reflection coated window (every packet passes, undeviated - the handler
    moves the packet into the surrounding medium itself and returns MC_REFRACTED), a ring
    up to ``r_black`` is a black absorber (the packet is reflected with zero weight), the rest
    of the surface is a glass of refractive index ``n_glass`` (n2 / cc override, Fresnel by
    the kernel)."""
    cltypes = _cltypes(mc)

    class UserWindow(mc.mcsurface.SurfaceLayoutAny):
        def cl_type(self, mc_):
            T = mc_.types
            class ClUserWindow(cltypes.Structure):
                _fields_ = [('r2_window', T.mc_fp_t), ('r2_black', T.mc_fp_t),
                            ('n_glass', T.mc_fp_t)]
            return ClUserWindow

        def cl_declaration(self, mc_):
            return 'struct MC_STRUCT_ATTRIBUTES Mc{}SurfaceLayout{{ mc_fp_t r2_window; ' \
                   'mc_fp_t r2_black; mc_fp_t n_glass; }};\n'.format(self.location.capitalize())

        def cl_implementation(self, mc_):
            loc = self.location
            Loc = loc.capitalize()
            outside = 'mcsim_top_layer_index(mcsim)' if loc == 'top' else \
                'mcsim_bottom_layer_index(mcsim)'
            return '''
void dbg_print_{loc}_surface_layout(__mc_surface_mem const Mc{Loc}SurfaceLayout *layout){{
	dbg_print("user-written window:");
	dbg_print_float(INDENT "n_glass:", layout->n_glass);
}};

inline int mcsim_{loc}_surface_layout_handler(McSim *mcsim, mc_fp_t *n2, mc_fp_t *cc){{
	__mc_surface_mem const struct Mc{Loc}SurfaceLayout *layout = mcsim_{loc}_surface_layout(mcsim);
	mc_fp_t r2 = mcsim_position_r2(mcsim);

	if (r2 <= layout->r2_window){{
		mcsim_set_current_layer_index(mcsim, {outside});
		return MC_REFRACTED;
	}};
	if (r2 <= layout->r2_black){{
		mcsim_reverse_direction_z(mcsim);
		mcsim_set_weight(mcsim, FP_0);
		return MC_REFLECTED;
	}};
	*n2 = layout->n_glass;
	*cc = cos_critical(mc_layer_n(mcsim_current_layer(mcsim)), layout->n_glass);
	return MC_SURFACE_LAYOUT_CONTINUE;
}};
'''.format(loc=loc, Loc=Loc, outside=outside)

        def __init__(self, r_window, r_black, n_glass):
            super().__init__()
            self._rw, self._rb, self._n = float(r_window), float(r_black), float(n_glass)

        def cl_pack(self, mc_, target=None):
            if target is None:
                target = self.cl_type(mc_)()
            target.r2_window = self._rw**2
            target.r2_black = self._rb**2
            target.n_glass = self._n
            return target

        def todict(self):
            return {'type': 'UserWindow', 'r_window': self._rw, 'r_black': self._rb,
                    'n_glass': self._n}

    return UserWindow


def user_window(mc, r_window, r_black, n_glass):
    return _user_window_class(mc)(r_window, r_black, n_glass)


# ---- Rayleigh: the reference's own class with its fragment repaired ----------------------------
@functools.lru_cache(maxsize=None)
def _fixed_rayleigh_class(mc):
    """``xopto.mcbase.mcpf.Rayleigh`` packs correctly, but its OpenCL-C text does not compile
    (rayleigh.py:96 lacks a ``;``, :99 closes a block with ``);``, and the constant FP_1d27 it
    uses expands to a literal with an ``ff`` suffix, mcbase.template.h:484).  This subclass -
    used only to produce golden vectors with the reference's kernel - carries the same text
    with those slips repaired."""
    class FixedRayleigh(mc.mcpf.Rayleigh):
        @staticmethod
        def cl_implementation(mc_):
            return '''
void dbg_print_pf(const McPf *pf) {
	dbg_print("Rayleigh scattering phase function:");
	dbg_print_float(INDENT "gamma:", pf->gamma);
};

inline mc_fp_t mcsim_pf_sample_angles(McSim *mcsim, mc_fp_t *azimuth){
	mc_fp_t tmp, cos_theta;
	mc_fp_t gamma = mcsim_current_pf(mcsim)->gamma;
	mc_fp_t a = mcsim_current_pf(mcsim)->a;
	mc_fp_t b = mcsim_current_pf(mcsim)->b;

	*azimuth = FP_2PI*mcsim_random(mcsim);

	if (gamma == FP_1) {
		cos_theta = FP_2*mcsim_random(mcsim) - FP_1;
	} else {
		b = b*(FP_1 - FP_2*mcsim_random(mcsim));
		tmp = mc_sqrt(b*b*FP_0p25 + a*a*a*FP_LITERAL(0.037037037037037035));
		cos_theta = mc_cbrt(-FP_0p5*b + tmp) + mc_cbrt(-FP_0p5*b - tmp);
	};

	return mc_fclip(cos_theta, -FP_1, FP_1);
};
'''
    # (exported under the reference's name: to_dict() / adoption see a `Rayleigh`)
    FixedRayleigh.__name__ = FixedRayleigh.__qualname__ = 'Rayleigh'
    return FixedRayleigh


def rayleigh(mc, gamma):
    """Rayleigh phase function: the engine's class, or - for the reference package - the
    reference's class with its fragment repaired."""
    if mc.__name__.startswith('xopto'):
        return _fixed_rayleigh_class(mc)(gamma)
    return mc.mcpf.Rayleigh(gamma)


# ---- trace (mcbase/mctrace.py:541-585) ------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _user_trace_class(mc, squared: bool):
    """A Trace whose device side is written by the user.  ``squared=False`` restates the
    built-in event record (results must equal those of ``Trace`` bit for bit);
    ``squared=True`` records the *square* of the packet weight in field 6 - something the
    reference does not ship.  The host side (buffers, result arrays, filter) is inherited."""
    class UserTrace(mc.mctrace.Trace):
        @staticmethod
        def cl_declaration(mc_):
            return 'struct MC_STRUCT_ATTRIBUTES McTrace{ mc_int_t max_events; ' \
                   'mc_size_t data_buffer_offset; mc_size_t count_buffer_offset; ' \
                   'mc_uint_t event_mask; };\n'

        @staticmethod
        def cl_implementation(mc_):
            return '''
void dbg_print_trace(__mc_trace_mem const McTrace *trace){
	dbg_print("user-written trace:");
	dbg_print_int(INDENT "max_events:", trace->max_events);
};

inline int mcsim_trace_event(McSim *mcsim, mc_uint_t event_count){
	__mc_trace_mem const McTrace *trace = mcsim_trace(mcsim);
	mc_size_t pos = mc_min(event_count, trace->max_events - 1)*TRACE_ENTRY_LEN +
		mcsim_packet_index(mcsim)*trace->max_events*TRACE_ENTRY_LEN +
		trace->data_buffer_offset;

	#if MC_USE_EVENTS
		if (!(trace->event_mask & mcsim_event_flags(mcsim)))
			return 0;
	#endif

	mcsim_float_buffer(mcsim)[pos++] = mcsim_position_x(mcsim);
	mcsim_float_buffer(mcsim)[pos++] = mcsim_position_y(mcsim);
	mcsim_float_buffer(mcsim)[pos++] = mcsim_position_z(mcsim);
	mcsim_float_buffer(mcsim)[pos++] = mcsim_direction_x(mcsim);
	mcsim_float_buffer(mcsim)[pos++] = mcsim_direction_y(mcsim);
	mcsim_float_buffer(mcsim)[pos++] = mcsim_direction_z(mcsim);
	mcsim_float_buffer(mcsim)[pos++] = %s;
	#if MC_TRACK_OPTICAL_PATHLENGTH
		mcsim_float_buffer(mcsim)[pos++] = mcsim_optical_pathlength(mcsim);
	#else
		mcsim_float_buffer(mcsim)[pos++] = FP_0;
	#endif
	return 1;
};

inline void mcsim_trace_complete(McSim *mcsim, mc_uint_t event_count){
	mcsim_integer_buffer(mcsim)[
		mcsim_trace(mcsim)->count_buffer_offset +
		mcsim_packet_index(mcsim)] = (mc_int_t)event_count;
};
''' % ('mcsim_weight(mcsim)*mcsim_weight(mcsim)' if squared else 'mcsim_weight(mcsim)')

    return UserTrace


def user_trace(mc, squared=False, **kw):
    return _user_trace_class(mc, bool(squared))(**kw)


# ---- voxel geometry: a user-written source (mcvox/mcsource) -----------------------------------
@functools.lru_cache(maxsize=None)
def _user_vox_beam_class(mc):
    """Collimated beam that starts INSIDE the voxel box at ``position`` along ``direction``
    with a weight that depends on the refractive index of the voxel it starts in (so the
    fragment uses the voxel / material accessors of the voxelised simulator)."""
    cltypes = _cltypes(mc)

    class UserVoxBeam(mc.mcsource.Source):
        @staticmethod
        def cl_type(mc_):
            T = mc_.types
            class ClUserVoxBeam(cltypes.Structure):
                _fields_ = [('position', T.mc_point3f_t), ('direction', T.mc_point3f_t)]
            return ClUserVoxBeam

        @staticmethod
        def cl_declaration(mc_):
            return 'struct MC_STRUCT_ATTRIBUTES McSource{ mc_point3f_t position; ' \
                   'mc_point3f_t direction; };\n'

        @staticmethod
        def cl_implementation(mc_):
            return '''
void dbg_print_source(__mc_source_mem const McSource *src){
	dbg_print("user-written beam inside the voxel box:");
	dbg_print_point3f(INDENT "position:", &src->position);
};

inline void mcsim_launch(McSim *mcsim){
	__mc_source_mem const McSource *src = mcsim_source(mcsim);
	mc_point3_t voxel;
	mc_fp_t n_here, n_out;

	mcsim_set_position(mcsim, &src->position);
	mcsim_set_direction(mcsim, &src->direction);
	voxel.x = mc_int(mc_fdiv(src->position.x - mcsim_top_left_x(mcsim), mcsim_voxel_size_x(mcsim)));
	voxel.y = mc_int(mc_fdiv(src->position.y - mcsim_top_left_y(mcsim), mcsim_voxel_size_y(mcsim)));
	voxel.z = mc_int(mc_fdiv(src->position.z - mcsim_top_left_z(mcsim), mcsim_voxel_size_z(mcsim)));
	voxel.x = mc_clip(voxel.x, 0, mcsim_shape_x(mcsim) - 1);
	voxel.y = mc_clip(voxel.y, 0, mcsim_shape_y(mcsim) - 1);
	voxel.z = mc_clip(voxel.z, 0, mcsim_shape_z(mcsim) - 1);
	mcsim_set_voxel_index(mcsim, &voxel);
	/* as if the beam had entered from the surrounding medium at normal incidence */
	n_here = mc_material_n(mcsim_voxel_material(mcsim, &voxel));
	n_out = mc_material_n(mcsim_surrounding_material(mcsim));
	mcsim_set_weight(mcsim, FP_1 - mc_fdiv((n_out - n_here)*(n_out - n_here),
		(n_out + n_here)*(n_out + n_here)));
};
'''

        def __init__(self, position, direction):
            super().__init__()
            self.position = np.asarray(position, dtype=np.float64)
            d = np.asarray(direction, dtype=np.float64)
            self.direction = d/np.linalg.norm(d)

        def cl_pack(self, mc_, target=None):
            if target is None:
                target = self.cl_type(mc_)()
            target.position.fromarray(self.position)
            target.direction.fromarray(self.direction)
            return target, None, None

        def todict(self):
            return {'type': 'UserVoxBeam', 'position': self.position.tolist(),
                    'direction': self.direction.tolist()}

    return UserVoxBeam


def user_vox_beam(mc, position, direction):
    return _user_vox_beam_class(mc)(position, direction)


# ---- cylindrical geometry: a user-written source (mccyl/mcsource) ------------------------------
@functools.lru_cache(maxsize=None)
def _user_cyl_beam_class(mc):
    """Collimated beam that starts INSIDE the cylinder stack at ``position`` along
    ``direction``; the fragment finds the layer of the start point with the layer accessors
    of the cylindrical simulator."""
    cltypes = _cltypes(mc)

    class UserCylBeam(mc.mcsource.Source):
        @staticmethod
        def cl_type(mc_):
            T = mc_.types
            class ClUserCylBeam(cltypes.Structure):
                _fields_ = [('position', T.mc_point3f_t), ('direction', T.mc_point3f_t)]
            return ClUserCylBeam

        @staticmethod
        def cl_declaration(mc_):
            return 'struct MC_STRUCT_ATTRIBUTES McSource{ mc_point3f_t position; ' \
                   'mc_point3f_t direction; };\n'

        @staticmethod
        def cl_implementation(mc_):
            return '''
void dbg_print_source(__mc_source_mem const McSource *src){
	dbg_print("user-written beam inside the cylinders:");
	dbg_print_point3f(INDENT "position:", &src->position);
};

inline void mcsim_launch(McSim *mcsim){
	__mc_source_mem const McSource *src = mcsim_source(mcsim);
	mc_fp_t r = mc_sqrt(src->position.x*src->position.x + src->position.y*src->position.y);
	mc_int_t index = 1, i;

	for (i = 1; i < mcsim_layer_count(mcsim); ++i)
		if (r < mc_layer_r_outer(mcsim_layer(mcsim, i)) && r >= mc_layer_r_inner(mcsim_layer(mcsim, i)))
			index = i;
	mcsim_set_position(mcsim, &src->position);
	mcsim_set_direction(mcsim, &src->direction);
	mcsim_set_weight(mcsim, FP_1);
	mcsim_set_current_layer_index(mcsim, index);
};
'''

        def __init__(self, position, direction):
            super().__init__()
            self.position = np.asarray(position, dtype=np.float64)
            d = np.asarray(direction, dtype=np.float64)
            self.direction = d/np.linalg.norm(d)

        def cl_pack(self, mc_, target=None):
            if target is None:
                target = self.cl_type(mc_)()
            target.position.fromarray(self.position)
            target.direction.fromarray(self.direction)
            return target, None, None

        def todict(self):
            return {'type': 'UserCylBeam', 'position': self.position.tolist(),
                    'direction': self.direction.tolist()}

    return UserCylBeam


def user_cyl_beam(mc, position, direction):
    return _user_cyl_beam_class(mc)(position, direction)
