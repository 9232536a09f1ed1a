"""A tiny stand-in for an R session: loads the glue built against rstub/ and lets tests issue
.Call()s the way kmer_hash.R does (kmer_hash.R:5-28 in the reference)."""
import ctypes as C
import os

import numpy as np

from conftest import ROOT

RGLUE = os.path.join(ROOT, "kmer_hasher_b200", "rglue")
NILSXP, INTSXP, STRSXP, VECSXP, EXTPTRSXP = 0, 13, 16, 19, 22


class RError(RuntimeError):
    pass


class RSession:
    def __init__(self):
        self.stub = C.CDLL(os.path.join(RGLUE, "librstub.so"), mode=C.RTLD_GLOBAL)
        self.glue = C.CDLL(os.path.join(RGLUE, "kmer_hash_stub.so"), mode=C.RTLD_GLOBAL)
        s = self.stub
        vp = C.c_void_p
        s.rstub_string_vector.restype = vp
        s.rstub_string_vector.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_long)]
        s.rstub_int_vector.restype = vp
        s.rstub_int_vector.argtypes = [C.c_int, C.POINTER(C.c_int)]
        s.rstub_call.restype = vp
        s.rstub_call.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp), C.c_char_p, C.c_int]
        for name, res, args in [("TYPEOF", C.c_int, [vp]), ("Rf_length", C.c_int, [vp]), ("INTEGER", C.POINTER(C.c_int), [vp]),
                                ("CHAR", C.c_char_p, [vp]), ("STRING_ELT", vp, [vp, C.c_ssize_t]),
                                ("VECTOR_ELT", vp, [vp, C.c_ssize_t]), ("Rf_getAttrib", vp, [vp, vp]),
                                ("rstub_nrow", C.c_int, [vp]), ("rstub_ncol", C.c_int, [vp]),
                                ("rstub_finalize", None, [vp]), ("rstub_release", None, [vp]),
                                ("rstub_protect_depth", C.c_int, []), ("rstub_live_objects", C.c_long, []),
                                ("rstub_transient_bytes", C.c_size_t, []), ("R_ExternalPtrAddr", vp, [vp]),
                                ("R_ExternalPtrTag", vp, [vp])]:
            fn = getattr(s, name)
            fn.restype, fn.argtypes = res, args
        self.char_addr = C.CFUNCTYPE(C.c_void_p, C.c_void_p)(("CHAR", s))     # CHAR() as a raw address
        self.nil = C.c_void_p.in_dll(s, "R_NilValue").value
        self.names_sym = C.c_void_p.in_dll(s, "R_NamesSymbol").value
        self.glue.R_init_kmer_hash(None)          # what dyn.load() does

    # -- value construction (as.character / as.integer in the R closures) --
    def character(self, *strings):
        bs = [x if isinstance(x, bytes) else (x.tobytes() if isinstance(x, np.ndarray) else x.encode("latin-1")) for x in strings]
        arr = (C.c_char_p * len(bs))(*bs)
        lens = (C.c_long * len(bs))(*[len(b) for b in bs])
        return self.stub.rstub_string_vector(len(bs), arr, lens)

    def integer(self, *vals):
        return self.stub.rstub_int_vector(len(vals), (C.c_int * len(vals))(*vals))

    def call(self, name, *args):
        a = (C.c_void_p * len(args))(*args)
        err = C.create_string_buffer(1024)
        out = self.stub.rstub_call(name.encode(), len(args), a, err, 1024)
        if not out:
            raise RError(err.value.decode())
        return out

    # -- value inspection --
    def typeof(self, x):
        return self.stub.TYPEOF(x)

    def ints(self, x):
        n = self.stub.Rf_length(x)
        return np.ctypeslib.as_array(self.stub.INTEGER(x), shape=(n,)).copy() if n else np.empty(0, np.int32)

    def strings(self, x):
        return [self.stub.CHAR(self.stub.STRING_ELT(x, i)).decode() for i in range(self.stub.Rf_length(x))]

    def list_names(self, x):
        return self.strings(self.stub.Rf_getAttrib(x, self.names_sym))

    def elt(self, x, i):
        return self.stub.VECTOR_ELT(x, i)

    def dims(self, x):
        return self.stub.rstub_nrow(x), self.stub.rstub_ncol(x)

    # -- the R closures of kmer_hash.R, in Python --
    def make_kmer_hash(self, seq, k, do_sort=False):
        return self.call("make_kmer_h_index", self.character(seq), self.integer(k), self.integer(int(do_sort)))

    def kmer_pos(self, ptr, flag):
        r = self.call("kmer_positions", ptr, self.integer(flag))
        out = {}
        for i, n in enumerate(self.list_names(r)):
            e = self.elt(r, i)
            if e == self.nil:
                out[n] = None
            elif self.typeof(e) == STRSXP:
                out[n] = self.strings(e)
            else:
                v = self.ints(e)
                nr, nc = self.dims(e)
                out[n] = v.reshape(nc, nr) if nr > 0 else v      # t() of the column-major matrix
        self.stub.rstub_release(r)
        return out

    def kmer_pairs(self, ptr_a, ptr_b):
        """kmer.pairs (kmer_hash.R:30-34): t() of the 2 x M matrix .Call("kmer_pair_pos") returns."""
        r = self.call("kmer_pair_pos", ptr_a, ptr_b)
        nr, nc = self.dims(r)
        v = self.ints(r).reshape(nc, nr)
        self.stub.rstub_release(r)
        return v

    def seq_kmer_pos(self, ptr, seq, k):
        r = self.call("sequence_kmer_positions", ptr, self.character(seq), self.integer(k))
        nr, nc = self.dims(r)
        v = self.ints(r).reshape(nc, nr)
        self.stub.rstub_release(r)
        return v
