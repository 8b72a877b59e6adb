/* Host-side helper of the Python drop-in (not part of the C ABI): turns the reference's
 * `lidar_points` -- a Python sequence of separate (m_i, 2) float64 arrays, reference
 * src/dataloader.py:110-112 -- into the pointer and length arrays icpb_align_host_scans takes,
 * through the buffer protocol (about 50 ns per scan; a Python loop over `.ctypes.data` costs 1 us).
 * Built as _icpb_pyhelper.so and loaded with ctypes.PyDLL (the GIL is held during the call).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

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
    for (; k < n; ++k) {
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
