/* Host-side helper of the Python drop-in (not part of the C ABI): turns the reference's
 * `lidar_points` -- a Python sequence of separate (m_i, 2) float64 arrays, reference
 * src/dataloader.py:110-112 -- into the pointer and length arrays icpb_align_host_scans takes.
 * numpy arrays are read through numpy's C API (a few ns per scan: type check, dtype, shape and strides
 * straight from the array object); anything else that exports a buffer goes through the buffer
 * protocol (~40 ns); a Python loop over `.ctypes.data` costs 1 us per scan.
 * Built as _icpb_pyhelper.so and loaded with ctypes.PyDLL (the GIL is held during the call).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#ifdef ICPB_WITH_NUMPY
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
static int numpy_ready = 0;                       /* 0 not tried, 1 ok, -1 unavailable */
static int ensure_numpy(void)
{
    if (numpy_ready == 0) {
        numpy_ready = _import_array() == 0 ? 1 : -1;
        if (numpy_ready < 0) PyErr_Clear();
    }
    return numpy_ready > 0;
}
#endif

/* Fills ptrs[k] / lens[k] for every element that is a C-ordered (m, 2) float64 buffer.
 * Returns n when every element conformed, otherwise the index of the first element that did not
 * (the caller converts from there on), or -1 when `seq` is not a sequence of n elements. */
long long icpb_py_scan_ptrs(PyObject *seq, uint64_t *ptrs, int64_t *lens, long long n)
{
    PyObject *fast = PySequence_Fast(seq, "scans must be a sequence of arrays");
    if (!fast) { PyErr_Clear(); return -1; }
    if (PySequence_Fast_GET_SIZE(fast) != n) { Py_DECREF(fast); return -1; }
    PyObject **items = PySequence_Fast_ITEMS(fast);
    long long k = 0;
#ifdef ICPB_WITH_NUMPY
    const int have_numpy = ensure_numpy();
#endif
    for (; k < n; ++k) {
#ifdef ICPB_WITH_NUMPY
        if (have_numpy && PyArray_CheckExact(items[k])) {
            PyArrayObject *a = (PyArrayObject *)items[k];
            if (PyArray_TYPE(a) != NPY_DOUBLE || PyArray_NDIM(a) != 2 || !PyArray_ISNOTSWAPPED(a)) break;
            const npy_intp *sh = PyArray_DIMS(a), *st = PyArray_STRIDES(a);
            if (sh[1] != 2 || !(sh[0] == 0 || (st[1] == 8 && st[0] == 16))) break;
            if (((uintptr_t)PyArray_DATA(a) & 7u) != 0) break;
            ptrs[k] = (uint64_t)(uintptr_t)PyArray_DATA(a); lens[k] = (int64_t)sh[0];
            continue;
        }
#endif
        Py_buffer view;
        if (PyObject_GetBuffer(items[k], &view, PyBUF_STRIDES | PyBUF_FORMAT) != 0) { PyErr_Clear(); break; }
        const char *f = view.format ? view.format : "B";
        if (*f == '<' || *f == '=' || *f == '@') ++f;
        const int ok = view.ndim == 2 && view.itemsize == 8 && f[0] == 'd' && f[1] == '\0' && view.shape[1] == 2 &&
                       (view.shape[0] == 0 || (view.strides[1] == 8 && view.strides[0] == 16));
        if (ok) { ptrs[k] = (uint64_t)(uintptr_t)view.buf; lens[k] = (int64_t)view.shape[0]; }
        PyBuffer_Release(&view);
        if (!ok) break;
    }
    Py_DECREF(fast);
    return k;
}
