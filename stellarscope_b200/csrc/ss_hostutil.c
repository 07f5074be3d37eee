/* Host-side helper of the Python boundary (CPython C API, plain C).
 *
 * The reference hands the pooling driver its rows as Python containers:
 * ``st.bcode_ridx_map`` = dict barcode -> set of row ids (model.py:875, 1053) and
 * builds one N x 1 selector per model from them (row_identity_matrix,
 * sparse_plus.py:408-441; model.py:722-723, 742-743).  The device layout wants
 * the inverse map, row -> model id, as one int32 array.  Walking a million
 * Python ints through the iterator protocol costs more than the whole fit on
 * the GPU, so this module walks the sets' hash tables directly.
 *
 *   fill_row_segments(groups, ids, seg) -> number of rows assigned
 *     groups  sequence of sets (any iterable of ints is accepted)
 *     ids     int32 buffer, ids[i] = model id of groups[i]
 *     seg     writable int32 buffer [N], pre-filled with -1
 *   Raises ValueError when a row is claimed by two different models, IndexError
 *   when a row id is outside [0, N).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

/* The fast paths below read the hash tables of exact sets (_PySet_NextEntry) and the single digit of compact ints
 * (PyUnstable_Long_IsCompact / CompactValue).  Both exist in CPython 3.12 (the interpreter of this image); 3.13
 * no longer declares the first in its public headers, earlier versions lack the second.  On any other version the
 * module still builds: the set walks go through the iterator protocol and PyLong_AsSsize_t (slower, same results). */
#if PY_VERSION_HEX >= 0x030C0000 && PY_VERSION_HEX < 0x030D0000
#define SS_FAST_SETS 1
#else
#define SS_FAST_SETS 0
#endif

/* ---- threaded fast path ---------------------------------------------------------
 * Exact sets of exact, compact (single-digit) ints -- what the reference's loader
 * produces -- are walked by a few worker threads while the calling thread keeps the
 * GIL: the workers only READ the set tables and the int objects (no reference
 * counts, no allocation, no exceptions).  Anything else makes a worker bail out and
 * the serial, fully checked path below redoes the job.  Rows are claimed with a
 * compare-and-swap so that a row claimed by two models is detected whatever the
 * interleaving. */
typedef struct {
  PyObject** items;
  const int32_t* ids;
  Py_ssize_t g0, g1, n;
  int32_t* seg;
  Py_ssize_t count;
  int status; /* 0 ok, 1 unsupported object -> serial path, 2 conflict / range -> serial path reports it */
} worker_t;

static void* worker_main(void* arg) {
  worker_t* w = (worker_t*)arg;
#if !SS_FAST_SETS
  w->status = 1;                 /* no public way to walk a set without the GIL: the serial path does the job */
  return NULL;
#else
  Py_ssize_t count = 0;
  for (Py_ssize_t g = w->g0; g < w->g1; ++g) {
    PyObject* s = w->items[g];
    const int32_t id = w->ids[g];
    Py_ssize_t pos = 0;
    PyObject* key;
    Py_hash_t hash;
    while (_PySet_NextEntry(s, &pos, &key, &hash)) {
      if (!PyLong_CheckExact(key) || !PyUnstable_Long_IsCompact((PyLongObject*)key)) {
        w->status = 1;
        return NULL;
      }
      const Py_ssize_t row = PyUnstable_Long_CompactValue((PyLongObject*)key);
      if (row < 0 || row >= w->n) {
        w->status = 2;
        return NULL;
      }
      int32_t expect = -1;
      if (__atomic_compare_exchange_n(&w->seg[row], &expect, id, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
        ++count;
      } else if (expect != id) {
        w->status = 2;
        return NULL;
      }
    }
  }
  w->count = count;
  return NULL;
#endif
}

/* returns the number of rows assigned, -1 when the serial path has to take over (seg is then
 * partially written and must be reset by the caller) */
static Py_ssize_t fill_threaded(PyObject** items, Py_ssize_t ng, const int32_t* ids, int32_t* seg, Py_ssize_t n) {
#if !SS_FAST_SETS
  return -1;
#endif
  Py_ssize_t total = 0;
  for (Py_ssize_t g = 0; g < ng; ++g) {
    if (!PyAnySet_CheckExact(items[g])) return -1;
    total += PySet_GET_SIZE(items[g]);
  }
  long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
  int nt = (int)(ncpu < 1 ? 1 : (ncpu > 16 ? 16 : ncpu));
  { const char* e = getenv("SS_HOST_THREADS"); if (e) nt = atoi(e) < 1 ? 1 : (atoi(e) > 16 ? 16 : atoi(e)); }
  if (total < 200000 || nt < 2) return -1;  /* small inputs: the serial path is as fast */
  worker_t w[16];
  pthread_t th[16];
  Py_ssize_t g = 0, acc = 0;
  for (int i = 0; i < nt; ++i) {
    w[i].items = items; w[i].ids = ids; w[i].seg = seg; w[i].n = n; w[i].count = 0; w[i].status = 0;
    w[i].g0 = g;
    const Py_ssize_t target = total * (i + 1) / nt;
    while (g < ng && (acc < target || i == nt - 1)) acc += PySet_GET_SIZE(items[g++]);
    w[i].g1 = g;
  }
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int started = 0;
  for (int i = 0; i < nt; ++i) {
    if (pthread_create(&th[i], NULL, worker_main, &w[i]) != 0) break;
    ++started;
  }
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (getenv("SS_HOST_DEBUG")) fprintf(stderr, "threads: %.3f ms\n", 1e3 * ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)));
  if (getenv("SS_HOST_DEBUG")) for (int i = 0; i < nt; ++i) fprintf(stderr, "worker %d: groups %zd..%zd count %zd status %d\n", i, w[i].g0, w[i].g1, w[i].count, w[i].status);
  if (started < nt) return -1;
  Py_ssize_t count = 0;
  for (int i = 0; i < nt; ++i) {
    if (w[i].status != 0) return -1;
    count += w[i].count;
  }
  return count;
}

static int assign_row(PyObject* key, int32_t id, int32_t* seg, Py_ssize_t n, Py_ssize_t* count) {
  Py_ssize_t row = PyLong_AsSsize_t(key);
  if (row == -1 && PyErr_Occurred()) {
    PyErr_Clear();
    /* numpy integer scalars etc. */
    PyObject* ix = PyNumber_Index(key);
    if (!ix) return -1;
    row = PyLong_AsSsize_t(ix);
    Py_DECREF(ix);
    if (row == -1 && PyErr_Occurred()) return -1;
  }
  if (row < 0 || row >= n) {
    PyErr_Format(PyExc_IndexError, "row id %zd outside [0, %zd)", row, n);
    return -1;
  }
  if (seg[row] != -1 && seg[row] != id) {
    PyErr_Format(PyExc_ValueError, "row %zd belongs to more than one model (%d and %d)", row, (int)seg[row], (int)id);
    return -1;
  }
  if (seg[row] == -1) ++*count;
  seg[row] = id;
  return 0;
}

static PyObject* fill_row_segments(PyObject* self, PyObject* args) {
  PyObject *groups, *ids_obj, *seg_obj;
  if (!PyArg_ParseTuple(args, "OOO", &groups, &ids_obj, &seg_obj)) return NULL;
  Py_buffer ids, seg;
  if (PyObject_GetBuffer(ids_obj, &ids, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) return NULL;
  if (PyObject_GetBuffer(seg_obj, &seg, PyBUF_C_CONTIGUOUS | PyBUF_WRITABLE | PyBUF_FORMAT) < 0) {
    PyBuffer_Release(&ids);
    return NULL;
  }
  PyObject* fast = NULL;
  PyObject* result = NULL;
  if (ids.itemsize != 4 || seg.itemsize != 4) {
    PyErr_SetString(PyExc_TypeError, "ids and seg must be int32 buffers");
    goto done;
  }
  fast = PySequence_Fast(groups, "groups must be a sequence");
  if (!fast) goto done;
  {
    const Py_ssize_t ng = PySequence_Fast_GET_SIZE(fast);
    const Py_ssize_t n = seg.len / 4;
    int32_t* segp = (int32_t*)seg.buf;
    const int32_t* idp = (const int32_t*)ids.buf;
    Py_ssize_t count = 0;
    if (ids.len / 4 < ng) {
      PyErr_SetString(PyExc_ValueError, "ids is shorter than groups");
      goto done;
    }
    {
      const Py_ssize_t c = fill_threaded(PySequence_Fast_ITEMS(fast), ng, idp, segp, n);
      if (c >= 0) {
        result = PyLong_FromSsize_t(c);
        goto done;
      }
      /* serial path: start again from a clean map (seg was pre-filled with -1 by contract) */
      for (Py_ssize_t i = 0; i < n; ++i) segp[i] = -1;
    }
    for (Py_ssize_t g = 0; g < ng; ++g) {
      PyObject* s = PySequence_Fast_GET_ITEM(fast, g);
      const int32_t id = idp[g];
#if SS_FAST_SETS
      if (PyAnySet_Check(s)) {
        Py_ssize_t pos = 0;
        PyObject* key;
        Py_hash_t hash;
        while (_PySet_NextEntry(s, &pos, &key, &hash)) {
          if (assign_row(key, id, segp, n, &count) < 0) goto done;
        }
      } else
#endif
      {
        PyObject* it = PyObject_GetIter(s);
        if (!it) goto done;
        PyObject* key;
        while ((key = PyIter_Next(it))) {
          const int rc = assign_row(key, id, segp, n, &count);
          Py_DECREF(key);
          if (rc < 0) {
            Py_DECREF(it);
            goto done;
          }
        }
        Py_DECREF(it);
        if (PyErr_Occurred()) goto done;
      }
    }
    result = PyLong_FromSsize_t(count);
  }
done:
  Py_XDECREF(fast);
  PyBuffer_Release(&ids);
  PyBuffer_Release(&seg);
  return result;
}

/* number_groups(groups, ids) -> n: ids[i] = running index of groups[i] among the NON-EMPTY containers
 * (-1 for an empty one: the reference skips barcodes without reads, model.py:735-739). */
static PyObject* number_groups(PyObject* self, PyObject* args) {
  PyObject *groups, *ids_obj;
  if (!PyArg_ParseTuple(args, "OO", &groups, &ids_obj)) return NULL;
  Py_buffer ids;
  if (PyObject_GetBuffer(ids_obj, &ids, PyBUF_C_CONTIGUOUS | PyBUF_WRITABLE | PyBUF_FORMAT) < 0) return NULL;
  PyObject* fast = PySequence_Fast(groups, "groups must be a sequence");
  PyObject* result = NULL;
  if (fast) {
    const Py_ssize_t ng = PySequence_Fast_GET_SIZE(fast);
    if (ids.itemsize != 4 || ids.len / 4 < ng) {
      PyErr_SetString(PyExc_ValueError, "ids must be an int32 buffer as long as groups");
    } else {
      int32_t* idp = (int32_t*)ids.buf;
      int32_t n = 0;
      int ok = 1;
      for (Py_ssize_t g = 0; g < ng && ok; ++g) {
        PyObject* s = PySequence_Fast_GET_ITEM(fast, g);
        Py_ssize_t len = PyAnySet_Check(s) ? PySet_GET_SIZE(s) : PyObject_Length(s);
        if (len < 0) ok = 0;
        else idp[g] = len > 0 ? n++ : -1;
      }
      if (ok) result = PyLong_FromLong(n);
    }
    Py_DECREF(fast);
  }
  PyBuffer_Release(&ids);
  return result;
}

/* pair_ids(read_index, bcode_map, umi_map, row_group) -> list of (barcode, umi) pairs in first-seen order
 * row_group[read_index[q]] = index of (bcode_map[q], umi_map[q]) in that list, for every read q in the
 * iteration order of read_index -- the integer coding of the reference's bcumi_read dict
 * (Stellarscope.dedup_umi, model.py:1161-1164). */
static PyObject* pair_ids(PyObject* self, PyObject* args) {
  PyObject *read_index, *bmap, *umap, *rg_obj;
  if (!PyArg_ParseTuple(args, "O!O!O!O", &PyDict_Type, &read_index, &PyDict_Type, &bmap, &PyDict_Type, &umap, &rg_obj))
    return NULL;
  Py_buffer rg;
  if (PyObject_GetBuffer(rg_obj, &rg, PyBUF_C_CONTIGUOUS | PyBUF_WRITABLE | PyBUF_FORMAT) < 0) return NULL;
  PyObject* ids = NULL;
  PyObject* result = NULL;
  if (rg.itemsize != 4) {
    PyErr_SetString(PyExc_TypeError, "row_group must be an int32 buffer");
    goto done;
  }
  ids = PyDict_New();
  if (!ids) goto done;
  {
    int32_t* out = (int32_t*)rg.buf;
    const Py_ssize_t n = rg.len / 4;
    Py_ssize_t pos = 0;
    PyObject *qname, *rowobj;
    while (PyDict_Next(read_index, &pos, &qname, &rowobj)) {
      const Py_ssize_t row = PyLong_AsSsize_t(rowobj);
      if (row == -1 && PyErr_Occurred()) goto done;
      if (row < 0 || row >= n) {
        PyErr_Format(PyExc_IndexError, "row id %zd outside [0, %zd)", row, n);
        goto done;
      }
      PyObject* bc = PyDict_GetItemWithError(bmap, qname);
      PyObject* um = bc ? PyDict_GetItemWithError(umap, qname) : NULL;
      if (!bc || !um) {
        if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, qname);
        goto done;
      }
      PyObject* key = PyTuple_Pack(2, bc, um);
      if (!key) goto done;
      PyObject* idobj = PyDict_GetItemWithError(ids, key);
      Py_ssize_t id;
      if (idobj) {
        id = PyLong_AsSsize_t(idobj);
      } else {
        if (PyErr_Occurred()) { Py_DECREF(key); goto done; }
        id = PyDict_GET_SIZE(ids);
        PyObject* v = PyLong_FromSsize_t(id);
        if (!v || PyDict_SetItem(ids, key, v) < 0) { Py_XDECREF(v); Py_DECREF(key); goto done; }
        Py_DECREF(v);
      }
      Py_DECREF(key);
      out[row] = (int32_t)id;
    }
    result = PyDict_Keys(ids);
  }
done:
  Py_XDECREF(ids);
  PyBuffer_Release(&rg);
  return result;
}

/* code_mappings(mappings, read_index, feat_index, read, feat, alnscore, alnlen)
 * mappings: list of (code, query_name, feat_name, alnscore, alnlen) tuples, the loader's output
 * (Stellarscope._load_sequential, model.py:1131-1132).  Integer-codes them for ss_scores_build:
 * read[i] = read_index[query_name], feat[i] = feat_index[feat_name] -- the two dict look-ups of
 * mapping_to_matrix (model.py:967-968), KeyError like there. */
static int get_i32_buffer(PyObject* obj, Py_buffer* b, Py_ssize_t n, const char* name) {
  if (PyObject_GetBuffer(obj, b, PyBUF_C_CONTIGUOUS | PyBUF_WRITABLE) < 0) return -1;
  if (b->itemsize != 4 || b->len / 4 < n) {
    PyErr_Format(PyExc_TypeError, "%s must be an int32 buffer of at least %zd items", name, n);
    PyBuffer_Release(b);
    return -1;
  }
  return 0;
}

static PyObject* code_mappings(PyObject* self, PyObject* args) {
  PyObject *maps, *read_index, *feat_index, *o_read, *o_feat, *o_as, *o_len;
  if (!PyArg_ParseTuple(args, "O!O!O!OOOO", &PyList_Type, &maps, &PyDict_Type, &read_index, &PyDict_Type, &feat_index,
                        &o_read, &o_feat, &o_as, &o_len))
    return NULL;
  const Py_ssize_t n = PyList_GET_SIZE(maps);
  Py_buffer b_read, b_feat, b_as, b_len;
  if (get_i32_buffer(o_read, &b_read, n, "read") < 0) return NULL;
  if (get_i32_buffer(o_feat, &b_feat, n, "feat") < 0) { PyBuffer_Release(&b_read); return NULL; }
  if (get_i32_buffer(o_as, &b_as, n, "alnscore") < 0) { PyBuffer_Release(&b_read); PyBuffer_Release(&b_feat); return NULL; }
  if (get_i32_buffer(o_len, &b_len, n, "alnlen") < 0) {
    PyBuffer_Release(&b_read); PyBuffer_Release(&b_feat); PyBuffer_Release(&b_as);
    return NULL;
  }
  int32_t *rd = (int32_t*)b_read.buf, *ft = (int32_t*)b_feat.buf, *as = (int32_t*)b_as.buf, *ln = (int32_t*)b_len.buf;
  PyObject* result = NULL;
  for (Py_ssize_t i = 0; i < n; ++i) {
    PyObject* m = PyList_GET_ITEM(maps, i);
    if (!PyTuple_Check(m) || PyTuple_GET_SIZE(m) != 5) {
      PyErr_SetString(PyExc_ValueError, "mapping must be a (code, query_name, feat_name, alnscore, alnlen) tuple");
      goto done;
    }
    PyObject* r = PyDict_GetItemWithError(read_index, PyTuple_GET_ITEM(m, 1));
    if (!r) { if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, PyTuple_GET_ITEM(m, 1)); goto done; }
    PyObject* f = PyDict_GetItemWithError(feat_index, PyTuple_GET_ITEM(m, 2));
    if (!f) { if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, PyTuple_GET_ITEM(m, 2)); goto done; }
    const long rv = PyLong_AsLong(r), fv = PyLong_AsLong(f);
    const long av = PyLong_AsLong(PyTuple_GET_ITEM(m, 3)), lv = PyLong_AsLong(PyTuple_GET_ITEM(m, 4));
    if (PyErr_Occurred()) goto done;
    if (rv < 0 || rv > INT32_MAX || fv < 0 || fv > INT32_MAX || av < INT32_MIN / 2 || av > INT32_MAX / 2 || lv < INT32_MIN / 2 ||
        lv > INT32_MAX / 2) {
      PyErr_SetString(PyExc_OverflowError, "mapping field does not fit in int32");
      goto done;
    }
    rd[i] = (int32_t)rv;
    ft[i] = (int32_t)fv;
    as[i] = (int32_t)av;
    ln[i] = (int32_t)lv;
  }
  result = PyLong_FromSsize_t(n);
done:
  PyBuffer_Release(&b_read);
  PyBuffer_Release(&b_feat);
  PyBuffer_Release(&b_as);
  PyBuffer_Release(&b_len);
  return result;
}

/* row_labels(read_index, label_of_read, position_of_label, out, ids_or_None)
 * out[read_index[q]] = position_of_label.get(label_of_read[q], -1) for every read q -- the row -> output column
 * map of output_report (model.py:1389-1394: rows in read_index order, barcodes looked up per read name).
 * With `ids` (a dict, filled here) the labels are numbered in the order of the rows instead (label -> running id;
 * position_of_label is ignored): the UMI ids of agg_bc_umi.  KeyError like the reference when a read has no label. */
static PyObject* row_labels(PyObject* self, PyObject* args) {
  PyObject *read_index, *label_of, *pos_of, *out_obj, *ids;
  if (!PyArg_ParseTuple(args, "O!O!OOO", &PyDict_Type, &read_index, &PyDict_Type, &label_of, &pos_of, &out_obj, &ids))
    return NULL;
  if (ids != Py_None && !PyDict_Check(ids)) {
    PyErr_SetString(PyExc_TypeError, "ids must be a dict or None");
    return NULL;
  }
  if (ids == Py_None && !PyDict_Check(pos_of)) {
    PyErr_SetString(PyExc_TypeError, "position_of_label must be a dict");
    return NULL;
  }
  Py_buffer ob;
  if (PyObject_GetBuffer(out_obj, &ob, PyBUF_C_CONTIGUOUS | PyBUF_WRITABLE) < 0) return NULL;
  PyObject* result = NULL;
  if (ob.itemsize != 4) {
    PyErr_SetString(PyExc_TypeError, "out must be an int32 buffer");
    goto done;
  }
  {
    int32_t* out = (int32_t*)ob.buf;
    const Py_ssize_t n = ob.len / 4;
    const Py_ssize_t n_reads = PyDict_GET_SIZE(read_index);
    /* by row: the running ids of the UMI variant follow the row order, not the dict order */
    PyObject** by_row = NULL;
    if (ids != Py_None) {
      by_row = (PyObject**)calloc((size_t)(n > 0 ? n : 1), sizeof(PyObject*));
      if (!by_row) { PyErr_NoMemory(); goto done; }
    }
    Py_ssize_t pos = 0;
    PyObject *qname, *rowobj;
    int ok = 1;
    /* store_read_info (model.py:1045-1056) inserts a new read into both dicts at the same moment, with the same
     * key object: as long as the two dicts yield identical key objects side by side, the label is taken from the
     * parallel walk and no string is hashed; the first difference switches to look-ups for the rest */
    Py_ssize_t pos2 = 0;
    int parallel = 1;
    while (PyDict_Next(read_index, &pos, &qname, &rowobj)) {
      const Py_ssize_t row = PyLong_AsSsize_t(rowobj);
      if (row == -1 && PyErr_Occurred()) { ok = 0; break; }
      if (row < 0 || row >= n) {
        PyErr_Format(PyExc_IndexError, "row id %zd outside [0, %zd)", row, n);
        ok = 0;
        break;
      }
      PyObject* lab = NULL;
      if (parallel) {
        PyObject *k2, *v2;
        if (PyDict_Next(label_of, &pos2, &k2, &v2) && k2 == qname) lab = v2;
        else parallel = 0;
      }
      if (!lab) lab = PyDict_GetItemWithError(label_of, qname);
      if (!lab) {
        if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, qname);
        ok = 0;
        break;
      }
      if (by_row) {
        by_row[row] = lab;                       /* borrowed; label_of keeps it alive */
      } else {
        PyObject* p = PyDict_GetItemWithError(pos_of, lab);
        if (!p && PyErr_Occurred()) { ok = 0; break; }
        long v = -1;
        if (p) {
          v = PyLong_AsLong(p);
          if (v == -1 && PyErr_Occurred()) { ok = 0; break; }
        }
        out[row] = (int32_t)v;
      }
    }
    if (ok && by_row) {
      for (Py_ssize_t r = 0; r < n && ok; ++r) {
        PyObject* lab = by_row[r];
        if (!lab) { out[r] = -1; continue; }
        PyObject* idobj = PyDict_GetItemWithError(ids, lab);
        Py_ssize_t id;
        if (idobj) {
          id = PyLong_AsSsize_t(idobj);
        } else {
          if (PyErr_Occurred()) { ok = 0; break; }
          id = PyDict_GET_SIZE(ids);
          PyObject* v = PyLong_FromSsize_t(id);
          if (!v || PyDict_SetItem(ids, lab, v) < 0) { Py_XDECREF(v); ok = 0; break; }
          Py_DECREF(v);
        }
        out[r] = (int32_t)id;
      }
    }
    free(by_row);
    if (ok) result = PyLong_FromSsize_t(n_reads);
  }
done:
  PyBuffer_Release(&ob);
  return result;
}

/* ---- gather_rows: row-compacted copy of a CSR matrix ---------------------------------------------
 * A rank of a sharded run fits only the rows of its own models: it cuts them out of the
 * host matrix before the upload.  numpy would need an index array as long as the result; here every row
 * (run of consecutive rows) is one memcpy, spread over a few threads without the GIL.
 *   row_extents(indptr, rows, out_ptr) -> total      out_ptr int64[n+1] = prefix sums of the rows' lengths
 *   gather_rows(indptr, indices, data, rows, out_ptr, out_idx, out_dat)
 * indptr / indices: int32 or int64 buffers; data: any item size; rows: int64, any order. */
static inline int64_t ld_ix(const char* p, Py_ssize_t size, int64_t i) {
  return size == 4 ? (int64_t)((const int32_t*)p)[i] : ((const int64_t*)p)[i];
}
typedef struct {
  const char *ip, *idx, *dat;
  Py_ssize_t ip_size, idx_size, dat_size;
  const int64_t *rows, *optr;
  char *oidx, *odat;
  Py_ssize_t r0, r1;
} gather_t;
static void* gather_main(void* arg) {
  gather_t* g = (gather_t*)arg;
  Py_ssize_t r = g->r0;
  while (r < g->r1) {
    Py_ssize_t e = r + 1;
    while (e < g->r1 && g->rows[e] == g->rows[e - 1] + 1) ++e;      /* run of consecutive rows: one copy */
    const int64_t a = ld_ix(g->ip, g->ip_size, g->rows[r]), b = ld_ix(g->ip, g->ip_size, g->rows[e - 1] + 1);
    const int64_t o = g->optr[r];
    memcpy(g->oidx + o * g->idx_size, g->idx + a * g->idx_size, (size_t)((b - a) * g->idx_size));
    memcpy(g->odat + o * g->dat_size, g->dat + a * g->dat_size, (size_t)((b - a) * g->dat_size));
    r = e;
  }
  return NULL;
}
static int host_threads(void) {
  long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
  int nt = (int)(ncpu < 1 ? 1 : (ncpu > 16 ? 16 : ncpu));
  const char* e = getenv("SS_HOST_THREADS");
  if (e) nt = atoi(e) < 1 ? 1 : (atoi(e) > 16 ? 16 : atoi(e));
  return nt;
}
typedef struct {
  const char* ip; Py_ssize_t ip_size, nrow;
  const int64_t* rows; int64_t* optr;
  Py_ssize_t r0, r1;
  int64_t total, base;
  int pass, bad;
} ext_t;
static void* ext_main(void* arg) {
  ext_t* w = (ext_t*)arg;
  int64_t acc = w->pass ? w->base : 0;
  for (Py_ssize_t i = w->r0; i < w->r1; ++i) {
    const int64_t r = w->rows[i];
    if (r < 0 || r >= w->nrow) { w->bad = 1; return NULL; }
    if (w->pass) w->optr[i] = acc;
    acc += ld_ix(w->ip, w->ip_size, r + 1) - ld_ix(w->ip, w->ip_size, r);
  }
  w->total = acc - (w->pass ? w->base : 0);
  return NULL;
}
static PyObject* row_extents(PyObject* self, PyObject* args) {
  Py_buffer ip, rows, optr;
  if (!PyArg_ParseTuple(args, "y*y*w*", &ip, &rows, &optr)) return NULL;
  PyObject* result = NULL;
  const Py_ssize_t n = rows.len / 8, nrow = ip.len / ip.itemsize - 1;
  if ((ip.itemsize != 4 && ip.itemsize != 8) || rows.itemsize != 8 || optr.itemsize != 8 || optr.len / 8 != n + 1) {
    PyErr_SetString(PyExc_TypeError, "row_extents: indptr int32/int64, rows int64[n], out_ptr int64[n+1]");
    goto done;
  }
  {
    /* two passes over the rows on a few threads: chunk totals, then the prefix sums from the chunk bases */
    int nt = host_threads();
    if (n < 1000000) nt = 1;
    ext_t w[16];
    pthread_t th[16];
    int bad = 0;
    int64_t acc = 0;
    Py_BEGIN_ALLOW_THREADS
    for (int pass = 0; pass < 2; ++pass) {
      int64_t base = 0;
      for (int i = 0; i < nt; ++i) {
        if (pass == 0) {
          w[i].ip = (const char*)ip.buf; w[i].ip_size = ip.itemsize; w[i].nrow = nrow;
          w[i].rows = (const int64_t*)rows.buf; w[i].optr = (int64_t*)optr.buf;
          w[i].r0 = n * i / nt; w[i].r1 = n * (i + 1) / nt; w[i].bad = 0; w[i].total = 0;
        } else {
          w[i].base = base;
          base += w[i].total;
        }
        w[i].pass = pass;
      }
      int started = 0;
      for (int i = 1; i < nt; ++i) {
        if (pthread_create(&th[i], NULL, ext_main, &w[i]) != 0) break;
        ++started;
      }
      ext_main(&w[0]);
      for (int i = 1; i <= started; ++i) pthread_join(th[i], NULL);
      for (int i = started + 1; i < nt; ++i) ext_main(&w[i]);
      for (int i = 0; i < nt; ++i) bad |= w[i].bad;
      if (bad) break;
      if (pass == 1) acc = base;
    }
    Py_END_ALLOW_THREADS
    if (bad) {
      PyErr_Format(PyExc_IndexError, "row id outside [0, %zd)", nrow);
      goto done;
    }
    ((int64_t*)optr.buf)[n] = acc;
    result = PyLong_FromLongLong(acc);
  }
done:
  PyBuffer_Release(&ip); PyBuffer_Release(&rows); PyBuffer_Release(&optr);
  return result;
}
/* gather_chunk(indptr, src, rows, out_ptr, e0, e1, dst): elements [e0, e1) of the row-compacted copy of `src`
 * (what gather_rows would produce) written to dst[0 .. e1-e0) -- the staging ring of the upload is filled straight
 * from the host matrix, the compacted copy never exists on the host.  Called from several Python threads at once
 * (the GIL is released). */
static PyObject* gather_chunk(PyObject* self, PyObject* args) {
  Py_buffer ip, src, rows, optr, dst;
  long long e0, e1;
  if (!PyArg_ParseTuple(args, "y*y*y*y*LLw*", &ip, &src, &rows, &optr, &e0, &e1, &dst)) return NULL;
  PyObject* result = NULL;
  const Py_ssize_t n = rows.len / 8, isz = src.itemsize;
  if ((ip.itemsize != 4 && ip.itemsize != 8) || rows.itemsize != 8 || optr.itemsize != 8 || optr.len / 8 != n + 1 ||
      e0 < 0 || e1 < e0 || (e1 - e0) * isz > dst.len) {
    PyErr_SetString(PyExc_TypeError, "gather_chunk: mismatching buffers");
    goto done;
  }
  {
    const int64_t* op = (const int64_t*)optr.buf;
    const int64_t* rw = (const int64_t*)rows.buf;
    if (e1 > op[n]) { PyErr_SetString(PyExc_ValueError, "gather_chunk: range beyond the last row"); goto done; }
    Py_BEGIN_ALLOW_THREADS
    /* first row that holds element e0: largest r with op[r] <= e0 and op[r+1] > e0 */
    Py_ssize_t lo = 0, hi = n;
    while (lo < hi) {
      const Py_ssize_t mid = (lo + hi) / 2;
      if (op[mid + 1] <= e0) lo = mid + 1; else hi = mid;
    }
    Py_ssize_t r = lo;
    int64_t e = e0;
    char* d = (char*)dst.buf;
    while (e < e1 && r < n) {
      /* run of consecutive rows starting at r */
      Py_ssize_t re = r + 1;
      while (re < n && rw[re] == rw[re - 1] + 1 && op[re] < e1) ++re;
      const int64_t src0 = ld_ix((const char*)ip.buf, ip.itemsize, rw[r]) + (e - op[r]);
      const int64_t run_end = op[re] < e1 ? op[re] : e1;
      memcpy(d, (const char*)src.buf + src0 * isz, (size_t)((run_end - e) * isz));
      d += (run_end - e) * isz;
      e = run_end;
      r = re;
    }
    Py_END_ALLOW_THREADS
    result = PyLong_FromLongLong(e1 - e0);
  }
done:
  PyBuffer_Release(&ip); PyBuffer_Release(&src); PyBuffer_Release(&rows); PyBuffer_Release(&optr); PyBuffer_Release(&dst);
  return result;
}
static PyObject* gather_rows(PyObject* self, PyObject* args) {
  Py_buffer ip, idx, dat, rows, optr, oidx, odat;
  if (!PyArg_ParseTuple(args, "y*y*y*y*y*w*w*", &ip, &idx, &dat, &rows, &optr, &oidx, &odat)) return NULL;
  PyObject* result = NULL;
  const Py_ssize_t n = rows.len / 8;
  if ((ip.itemsize != 4 && ip.itemsize != 8) || rows.itemsize != 8 || optr.itemsize != 8 || optr.len / 8 != n + 1 ||
      oidx.itemsize != idx.itemsize || odat.itemsize != dat.itemsize) {
    PyErr_SetString(PyExc_TypeError, "gather_rows: mismatching buffers");
    goto done;
  }
  {
    const int64_t total = ((const int64_t*)optr.buf)[n];
    if (oidx.len / oidx.itemsize < total || odat.len / odat.itemsize < total) {
      PyErr_SetString(PyExc_ValueError, "gather_rows: output buffers too small");
      goto done;
    }
    int nt = host_threads();
    if (total < 1000000) nt = 1;
    gather_t g[16];
    pthread_t th[16];
    const int64_t* op = (const int64_t*)optr.buf;
    Py_ssize_t r = 0;
    for (int i = 0; i < nt; ++i) {
      g[i].ip = (const char*)ip.buf; g[i].idx = (const char*)idx.buf; g[i].dat = (const char*)dat.buf;
      g[i].ip_size = ip.itemsize; g[i].idx_size = idx.itemsize; g[i].dat_size = dat.itemsize;
      g[i].rows = (const int64_t*)rows.buf; g[i].optr = op; g[i].oidx = (char*)oidx.buf; g[i].odat = (char*)odat.buf;
      g[i].r0 = r;
      const int64_t target = total / nt * (i + 1);
      while (r < n && (i == nt - 1 || op[r] < target)) ++r;
      g[i].r1 = r;
    }
    Py_BEGIN_ALLOW_THREADS
    int started = 0;
    for (int i = 1; i < nt; ++i) {
      if (pthread_create(&th[i], NULL, gather_main, &g[i]) != 0) break;
      ++started;
    }
    gather_main(&g[0]);
    for (int i = 1; i <= started; ++i) pthread_join(th[i], NULL);
    for (int i = started + 1; i < nt; ++i) gather_main(&g[i]);      /* threads that could not be started */
    Py_END_ALLOW_THREADS
    result = PyLong_FromLongLong(total);
  }
done:
  PyBuffer_Release(&ip); PyBuffer_Release(&idx); PyBuffer_Release(&dat); PyBuffer_Release(&rows);
  PyBuffer_Release(&optr); PyBuffer_Release(&oidx); PyBuffer_Release(&odat);
  return result;
}

/* ---- sharding plan of a pooling mode: alignments per model, LPT packing, the rows of one rank -------------
 * parallel.plan_local_rows in numpy makes ~10 passes over per-row arrays (1.4x10^8 rows at atlas scale: seconds
 * on every rank, more than the fit).  Three helpers, the per-row ones threaded and without the GIL:
 *   segment_alignments(indptr, row_segment, seg_A)        seg_A[s] = alignments of the rows of model s
 *   lpt_assign(seg_A, world, model_rank)                  longest-processing-time packing, ties by index
 *                                                         (the same assignment as parallel.lpt_shards)
 *   count_rows(row_segment, model_rank, rank, counts)     rows of `rank` per chunk of rows (counts int64[NCH])
 *   fill_rows(row_segment, model_rank, local_id, rank, offsets, rows, local_seg)
 * Rows with row_segment < 0 or model_rank < 0 (wide models) are never selected here (numpy path of the caller). */
#define SS_NCH 64
typedef struct {
  const char* ip; Py_ssize_t ip_size;
  const int32_t* seg; Py_ssize_t n, S;
  int64_t* part;                 /* [S] private partial sums */
  const int32_t *model_rank, *local_id;
  int rank;
  int64_t *counts, *rows; const int64_t* offsets; int32_t* lseg;
  Py_ssize_t c0, c1;             /* chunk range of this thread */
  int mode;
} plan_t;
static inline void chunk_bounds(Py_ssize_t n, int c, Py_ssize_t* a, Py_ssize_t* b) {
  *a = (Py_ssize_t)((__int128)n * c / SS_NCH);
  *b = (Py_ssize_t)((__int128)n * (c + 1) / SS_NCH);
}
static void* plan_main(void* arg) {
  plan_t* w = (plan_t*)arg;
  for (Py_ssize_t c = w->c0; c < w->c1; ++c) {
    Py_ssize_t a, b;
    chunk_bounds(w->n, (int)c, &a, &b);
    if (w->mode == 0) {
      for (Py_ssize_t r = a; r < b; ++r) {
        const int32_t s = w->seg[r];
        if (s >= 0 && s < w->S) w->part[s] += ld_ix(w->ip, w->ip_size, r + 1) - ld_ix(w->ip, w->ip_size, r);
      }
    } else if (w->mode == 3) {
      for (Py_ssize_t r = a; r < b; ++r) {
        const int32_t s = w->seg[r];
        if (s >= 0 && s < w->S) w->part[s] += 1;
      }
    } else if (w->mode == 1) {
      int64_t k = 0;
      for (Py_ssize_t r = a; r < b; ++r) {
        const int32_t s = w->seg[r];
        k += (s >= 0 && s < w->S && w->model_rank[s] == w->rank);
      }
      w->counts[c] = k;
    } else {
      int64_t o = w->offsets[c];
      for (Py_ssize_t r = a; r < b; ++r) {
        const int32_t s = w->seg[r];
        if (s >= 0 && s < w->S && w->model_rank[s] == w->rank) {
          w->rows[o] = r;
          w->lseg[o] = w->local_id[s];
          ++o;
        }
      }
    }
  }
  return NULL;
}
static void plan_run(plan_t* proto, int nt) {
  plan_t w[16];
  pthread_t th[16];
  if (nt > 16) nt = 16;
  if (nt < 1) nt = 1;
  for (int i = 0; i < nt; ++i) {
    w[i] = *proto;
    w[i].c0 = (Py_ssize_t)SS_NCH * i / nt;
    w[i].c1 = (Py_ssize_t)SS_NCH * (i + 1) / nt;
  }
  int started = 0;
  for (int i = 1; i < nt; ++i) {
    if (pthread_create(&th[i], NULL, plan_main, &w[i]) != 0) break;
    ++started;
  }
  plan_main(&w[0]);
  for (int i = 1; i <= started; ++i) pthread_join(th[i], NULL);
  for (int i = started + 1; i < nt; ++i) plan_main(&w[i]);
}
static PyObject* segment_alignments(PyObject* self, PyObject* args) {
  Py_buffer ip, seg, out;
  if (!PyArg_ParseTuple(args, "y*y*w*", &ip, &seg, &out)) return NULL;
  PyObject* result = NULL;
  const Py_ssize_t n = seg.len / 4, S = out.len / 8;
  /* indptr of length 0: count ROWS per model instead of alignments (group_sizes) */
  const int rows_only = ip.len == 0;
  if ((!rows_only && ((ip.itemsize != 4 && ip.itemsize != 8) || ip.len / ip.itemsize != n + 1)) || seg.itemsize != 4 ||
      out.itemsize != 8) {
    PyErr_SetString(PyExc_TypeError, "segment_alignments: indptr int32/int64 [N+1], row_segment int32 [N], out int64 [S]");
    goto done;
  }
  {
    int nt = host_threads();
    if (n < 2000000) nt = 1;
    int64_t* parts = (int64_t*)calloc((size_t)nt * (size_t)(S > 0 ? S : 1), sizeof(int64_t));
    if (!parts) { PyErr_NoMemory(); goto done; }
    /* one private accumulator array per thread: chunk ranges are split evenly over nt threads by plan_run, so
       thread i is identified by its first chunk */
    plan_t w[16];
    pthread_t th[16];
    Py_BEGIN_ALLOW_THREADS
    for (int i = 0; i < nt; ++i) {
      memset(&w[i], 0, sizeof(plan_t));
      w[i].ip = (const char*)ip.buf; w[i].ip_size = ip.itemsize; w[i].seg = (const int32_t*)seg.buf;
      w[i].n = n; w[i].S = S; w[i].part = parts + (size_t)i * (size_t)(S > 0 ? S : 1); w[i].mode = rows_only ? 3 : 0;
      w[i].c0 = (Py_ssize_t)SS_NCH * i / nt; w[i].c1 = (Py_ssize_t)SS_NCH * (i + 1) / nt;
    }
    int started = 0;
    for (int i = 1; i < nt; ++i) {
      if (pthread_create(&th[i], NULL, plan_main, &w[i]) != 0) break;
      ++started;
    }
    plan_main(&w[0]);
    for (int i = 1; i <= started; ++i) pthread_join(th[i], NULL);
    for (int i = started + 1; i < nt; ++i) plan_main(&w[i]);
    int64_t* o = (int64_t*)out.buf;
    for (Py_ssize_t s = 0; s < S; ++s) {
      int64_t a = 0;
      for (int i = 0; i < nt; ++i) a += parts[(size_t)i * (size_t)S + s];
      o[s] = a;
    }
    Py_END_ALLOW_THREADS
    free(parts);
    result = PyLong_FromSsize_t(S);
  }
done:
  PyBuffer_Release(&ip); PyBuffer_Release(&seg); PyBuffer_Release(&out);
  return result;
}
typedef struct { int64_t size; int32_t idx; } lpt_item;
static int lpt_cmp(const void* a, const void* b) {
  const lpt_item *x = (const lpt_item*)a, *y = (const lpt_item*)b;
  if (x->size != y->size) return x->size > y->size ? -1 : 1;       /* larger first */
  return x->idx < y->idx ? -1 : (x->idx > y->idx);                 /* ties by index */
}
static PyObject* lpt_assign(PyObject* self, PyObject* args) {
  Py_buffer sizes, out;
  int world;
  if (!PyArg_ParseTuple(args, "y*iw*", &sizes, &world, &out)) return NULL;
  PyObject* result = NULL;
  const Py_ssize_t S = sizes.len / 8;
  if (sizes.itemsize != 8 || out.itemsize != 4 || out.len / 4 != S || world < 1 || world > 4096) {
    PyErr_SetString(PyExc_TypeError, "lpt_assign: sizes int64 [S], world, model_rank int32 [S]");
    goto done;
  }
  {
    lpt_item* it = (lpt_item*)malloc((size_t)(S > 0 ? S : 1) * sizeof(lpt_item));
    int64_t* load = (int64_t*)calloc((size_t)world, sizeof(int64_t));
    if (!it || !load) { free(it); free(load); PyErr_NoMemory(); goto done; }
    const int64_t* sz = (const int64_t*)sizes.buf;
    int32_t* mr = (int32_t*)out.buf;
    for (Py_ssize_t s = 0; s < S; ++s) { it[s].size = sz[s]; it[s].idx = (int32_t)s; }
    qsort(it, (size_t)S, sizeof(lpt_item), lpt_cmp);
    for (Py_ssize_t k = 0; k < S; ++k) {
      int best = 0;
      for (int r = 1; r < world; ++r)
        if (load[r] < load[best]) best = r;                          /* smallest load, ties by rank */
      mr[it[k].idx] = best;
      load[best] += it[k].size;
    }
    free(it); free(load);
    result = PyLong_FromSsize_t(S);
  }
done:
  PyBuffer_Release(&sizes); PyBuffer_Release(&out);
  return result;
}
static PyObject* select_rows_impl(PyObject* args, int fill) {
  Py_buffer seg, mr, lid, cnt, rows, lseg;
  int rank;
  memset(&lid, 0, sizeof(lid)); memset(&rows, 0, sizeof(rows)); memset(&lseg, 0, sizeof(lseg));
  if (!fill) {
    if (!PyArg_ParseTuple(args, "y*y*iw*", &seg, &mr, &rank, &cnt)) return NULL;
  } else {
    if (!PyArg_ParseTuple(args, "y*y*y*iy*w*w*", &seg, &mr, &lid, &rank, &cnt, &rows, &lseg)) return NULL;
  }
  PyObject* result = NULL;
  const Py_ssize_t n = seg.len / 4, S = mr.len / 4;
  if (seg.itemsize != 4 || mr.itemsize != 4 || cnt.itemsize != 8 || cnt.len / 8 != SS_NCH ||
      (fill && (lid.itemsize != 4 || lid.len / 4 != S || rows.itemsize != 8 || lseg.itemsize != 4))) {
    PyErr_SetString(PyExc_TypeError, "count_rows / fill_rows: mismatching buffers");
    goto done;
  }
  {
    plan_t w;
    memset(&w, 0, sizeof(w));
    w.seg = (const int32_t*)seg.buf; w.n = n; w.S = S; w.model_rank = (const int32_t*)mr.buf; w.rank = rank;
    if (!fill) {
      w.mode = 1; w.counts = (int64_t*)cnt.buf;
    } else {
      const int64_t* off = (const int64_t*)cnt.buf;
      w.mode = 2; w.offsets = off; w.local_id = (const int32_t*)lid.buf; w.rows = (int64_t*)rows.buf; w.lseg = (int32_t*)lseg.buf;
    }
    int nt = host_threads();
    if (n < 2000000) nt = 1;
    Py_BEGIN_ALLOW_THREADS
    plan_run(&w, nt);
    Py_END_ALLOW_THREADS
    result = PyLong_FromSsize_t(n);
  }
done:
  PyBuffer_Release(&seg); PyBuffer_Release(&mr); PyBuffer_Release(&cnt);
  if (fill) { PyBuffer_Release(&lid); PyBuffer_Release(&rows); PyBuffer_Release(&lseg); }
  return result;
}
static PyObject* count_rows(PyObject* self, PyObject* args) { return select_rows_impl(args, 0); }
static PyObject* fill_rows(PyObject* self, PyObject* args) { return select_rows_impl(args, 1); }

/* fill_ranges(starts, ends, offsets, rows, local_seg): the rows of model i of a rank are [starts[i], ends[i]) (row -> model
 * map sorted by model: parallel.plan_local_rows); rows[offsets[i] + k] = starts[i] + k, local_seg[...] = i.  Threaded
 * over the models in blocks of equal row count. */
typedef struct { const int64_t *st, *en, *off; int64_t* rows; int32_t* lseg; Py_ssize_t m0, m1; } range_t;
static void* range_main(void* arg) {
  range_t* w = (range_t*)arg;
  for (Py_ssize_t i = w->m0; i < w->m1; ++i) {
    int64_t o = w->off[i];
    for (int64_t r = w->st[i]; r < w->en[i]; ++r, ++o) { w->rows[o] = r; w->lseg[o] = (int32_t)i; }
  }
  return NULL;
}
static PyObject* fill_ranges(PyObject* self, PyObject* args) {
  Py_buffer st, en, off, rows, lseg;
  if (!PyArg_ParseTuple(args, "y*y*y*w*w*", &st, &en, &off, &rows, &lseg)) return NULL;
  PyObject* result = NULL;
  const Py_ssize_t m = st.len / 8, n = rows.len / 8;
  if (st.itemsize != 8 || en.itemsize != 8 || off.itemsize != 8 || rows.itemsize != 8 || lseg.itemsize != 4 ||
      en.len / 8 != m || off.len / 8 != m + 1 || lseg.len / 4 != n || (m > 0 && ((const int64_t*)off.buf)[m] != n)) {
    PyErr_SetString(PyExc_TypeError, "fill_ranges: starts / ends int64 [m], offsets int64 [m + 1], rows int64 [n], local_seg int32 [n]");
    goto done;
  }
  {
    int nt = host_threads();
    if (n < 2000000) nt = 1;
    if (nt > 16) nt = 16;
    range_t w[16];
    pthread_t th[16];
    const int64_t* o = (const int64_t*)off.buf;
    Py_BEGIN_ALLOW_THREADS
    Py_ssize_t m0 = 0;
    for (int i = 0; i < nt; ++i) {
      /* models up to the one that holds row (i + 1) n / nt */
      const int64_t want = (int64_t)((__int128)n * (i + 1) / nt);
      Py_ssize_t m1 = m0;
      if (i == nt - 1) m1 = m;
      else while (m1 < m && o[m1 + 1] <= want) ++m1;
      w[i].st = (const int64_t*)st.buf; w[i].en = (const int64_t*)en.buf; w[i].off = o;
      w[i].rows = (int64_t*)rows.buf; w[i].lseg = (int32_t*)lseg.buf; w[i].m0 = m0; w[i].m1 = m1;
      m0 = m1;
    }
    int started = 0;
    for (int i = 1; i < nt; ++i) {
      if (pthread_create(&th[i], NULL, range_main, &w[i]) != 0) break;
      ++started;
    }
    range_main(&w[0]);
    for (int i = 1; i <= started; ++i) pthread_join(th[i], NULL);
    for (int i = started + 1; i < nt; ++i) range_main(&w[i]);
    Py_END_ALLOW_THREADS
    result = PyLong_FromSsize_t(n);
  }
done:
  PyBuffer_Release(&st); PyBuffer_Release(&en); PyBuffer_Release(&off); PyBuffer_Release(&rows); PyBuffer_Release(&lseg);
  return result;
}

static PyMethodDef methods[] = {
    {"segment_alignments", segment_alignments, METH_VARARGS, "segment_alignments(indptr, row_segment, seg_A): alignments per model"},
    {"lpt_assign", lpt_assign, METH_VARARGS, "lpt_assign(seg_A, world, model_rank): longest-processing-time packing"},
    {"count_rows", count_rows, METH_VARARGS, "count_rows(row_segment, model_rank, rank, counts): rows of a rank per row chunk"},
    {"fill_rows", fill_rows, METH_VARARGS, "fill_rows(row_segment, model_rank, local_id, rank, offsets, rows, local_seg)"},
    {"fill_ranges", fill_ranges, METH_VARARGS, "fill_ranges(starts, ends, offsets, rows, local_seg): rows of consecutive-row models"},
    {"gather_chunk", gather_chunk, METH_VARARGS,
     "gather_chunk(indptr, src, rows, out_ptr, e0, e1, dst): a range of the row-compacted copy of src, written to dst"},
    {"row_extents", row_extents, METH_VARARGS, "row_extents(indptr, rows, out_ptr): prefix sums of the lengths of the given rows"},
    {"gather_rows", gather_rows, METH_VARARGS,
     "gather_rows(indptr, indices, data, rows, out_ptr, out_idx, out_dat): row-compacted copy of a CSR matrix"},
    {"row_labels", row_labels, METH_VARARGS,
     "row_labels(read_index, label_of_read, position_of_label, out, ids): out[row] = position (or running id) of the read's label"},
    {"code_mappings", code_mappings, METH_VARARGS,
     "code_mappings(mappings, read_index, feat_index, read, feat, alnscore, alnlen): integer-code the loader's mapping tuples"},
    {"pair_ids", pair_ids, METH_VARARGS,
     "pair_ids(read_index, bcode_map, umi_map, row_group): integer-code the (barcode, UMI) pair of every read"},
    {"number_groups", number_groups, METH_VARARGS,
     "number_groups(groups, ids): ids[i] = index of groups[i] among the non-empty containers, -1 if empty"},
    {"fill_row_segments", fill_row_segments, METH_VARARGS,
     "fill_row_segments(groups, ids, seg): seg[row] = ids[i] for every row of groups[i]"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_ss_hostutil", "host helpers of stellarscope_b200", -1, methods};

PyMODINIT_FUNC PyInit__ss_hostutil(void) { return PyModule_Create(&moddef); }
