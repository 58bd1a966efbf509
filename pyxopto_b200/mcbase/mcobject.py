"""Plugin protocol (mirror of ``xopto/mcbase/mcobject.py:29-170``).

A plugin object provides a packed ctypes struct (``cl_type`` / ``cl_pack``),
compile-time options (``cl_options``) and - instead of the reference's OpenCL-C
text (``cl_declaration`` / ``cl_implementation``) - the name of the hand-written
CUDA struct that implements it in ``csrc/kernels`` (``cu_type``).  Objects that
still carry OpenCL-C fragments (user plugins) are compiled through
``csrc/kernels/xo_clcompat*.cuh`` (mcbase/mcsim.py `_user_fragments`).
"""


class McObject:
    cu_type = None          # name of the CUDA struct in csrc/kernels (xo::...)

    def fetch_cl_type(self, mc):
        t = self.cl_type
        return t(mc) if callable(t) else t

    def fetch_cl_options(self, mc):
        opts = getattr(self, 'cl_options', None)
        if opts is None:
            return []
        return (opts(mc) if callable(opts) else opts) or []

    def fetch_cu_type(self, mc):
        t = self.cu_type
        return t(mc) if callable(t) else t
