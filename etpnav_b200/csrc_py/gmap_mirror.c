/* Array-backed mirror of the reference's GraphMap objects (vlnce_baselines/models/graph_utils.py:133-250) for the stateful
 * map packer (etpnav_b200/packing.py: GmapPacker).  Host code against the CPython API, loaded with ctypes.PyDLL: one call
 * walks the dictionaries of every map of the batch (what Python does at ~100 us per map) and keeps, per map, flat arrays
 * that are updated INCREMENTALLY — the rules are those of packing._EnvMirror, which stays as the pure-Python twin the
 * tests compare this file with:
 *   - shortest_dist is rebuilt as a NEW dict by update_graph (:256-257): same object = no update since the last pack;
 *   - node ids are append-only; one new node adds its own row / column to the two all-pairs tables and makes only the
 *     pairs of older nodes that got a route through it be read again (none for a leaf); anything else re-reads the tables;
 *   - ghosts are compared by id sequence and per-ghost front count; ghost_aug_pos is re-read after an update;
 *   - embedding pointers are read (row_pointer: device / dtype / shape / contiguity / no-gradient check) for new nodes and
 *     for ghosts whose front count moved only.
 * No arithmetic on the values happens here (pure re-layout into the blobs etp_gmap_pack documents). */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  PyObject* gm_sd;      /* strong ref: the shortest_dist dict the tables were read from */
  PyObject* ix;         /* dict node id -> PyLong index */
  PyObject** nid; int n, cap;
  double* npos; int32_t* steps; double* sd; int32_t* spl;          /* [cap,3] [cap+1] [cap,cap] [cap,cap] */
  /* ghosts, in the order of GraphMap.ghost_pos; every array has a spare twin ("2") the next state is built in, so the rows
   * of ghosts that did not change are copied from the previous state instead of being read from the dictionaries again */
  PyObject **gids, **gids2; int g, gcap; int32_t *flens, *flens2, *fcum, *fcum2; double *gpos, *gpos2;
  int32_t *fidx, *fidx2; int nnz, nnzcap;
  /* image rows */
  int64_t* nptr; PyObject* nkeep; int nimg;                         /* node pointers [cap], list of kept tensors */
  int64_t *gptr, *gptr2; float *gwt, *gwt2; PyObject **gten, **gten2;   /* ghost pointers / weights / kept tensors */
  int img_g;                                                        /* number of ghosts the image rows are valid for, or -1 */
  int gten_n;                                                       /* entries of gten that hold a reference */
  int img_ok;
} Env;

static void env_clear(Env* s) {
  Py_CLEAR(s->gm_sd); Py_CLEAR(s->ix); Py_CLEAR(s->nkeep);
  for (int i = 0; i < s->n; ++i) Py_XDECREF(s->nid[i]);
  for (int i = 0; i < s->g; ++i) Py_XDECREF(s->gids[i]);
  for (int i = 0; i < s->gten_n; ++i) Py_XDECREF(s->gten[i]);
  free(s->nid); free(s->npos); free(s->steps); free(s->sd); free(s->spl); free(s->nptr);
  free(s->gids); free(s->flens); free(s->fcum); free(s->gpos); free(s->fidx); free(s->gptr); free(s->gwt); free(s->gten);
  free(s->gids2); free(s->flens2); free(s->fcum2); free(s->gpos2); free(s->fidx2); free(s->gptr2); free(s->gwt2); free(s->gten2);
  memset(s, 0, sizeof(*s));
}
static void env_destroy(PyObject* cap) {
  Env* s = (Env*)PyCapsule_GetPointer(cap, "etp.gmap_mirror");
  if (s) { env_clear(s); free(s); }
}
static int env_init(Env* s) {
  memset(s, 0, sizeof(*s));
  s->ix = PyDict_New(); s->nkeep = PyList_New(0);
  s->img_ok = 1;
  s->img_g = -1;
  return (s->ix && s->nkeep) ? 0 : -1;
}
static int grow_nodes(Env* s, int need) {
  if (need <= s->cap) return 0;
  int cap = s->cap ? s->cap : 16;
  while (cap < need) cap *= 2;
  PyObject** nid = (PyObject**)calloc(cap, sizeof(PyObject*));
  double* npos = (double*)calloc((size_t)cap * 3, 8);
  int32_t* steps = (int32_t*)calloc(cap + 1, 4);
  double* sd = (double*)calloc((size_t)cap * cap, 8);
  int32_t* spl = (int32_t*)calloc((size_t)cap * cap, 4);
  int64_t* nptr = (int64_t*)calloc(cap, 8);
  if (!nid || !npos || !steps || !sd || !spl || !nptr) { PyErr_NoMemory(); return -1; }
  for (int i = 0; i < s->n; ++i) {
    nid[i] = s->nid[i]; nptr[i] = s->nptr[i]; steps[i] = s->steps[i];
    memcpy(npos + 3 * i, s->npos + 3 * i, 24);
    memcpy(sd + (size_t)i * cap, s->sd + (size_t)i * s->cap, (size_t)s->n * 8);
    memcpy(spl + (size_t)i * cap, s->spl + (size_t)i * s->cap, (size_t)s->n * 4);
  }
  free(s->nid); free(s->npos); free(s->steps); free(s->sd); free(s->spl); free(s->nptr);
  s->nid = nid; s->npos = npos; s->steps = steps; s->sd = sd; s->spl = spl; s->nptr = nptr; s->cap = cap;
  return 0;
}
#define ETP_REGROW(ptr, type, count, keep)                                              \
  do {                                                                                  \
    type* fresh_ = (type*)calloc((size_t)(count), sizeof(type));                        \
    if (!fresh_) { PyErr_NoMemory(); return -1; }                                       \
    if ((keep) > 0 && (ptr)) memcpy(fresh_, (ptr), (size_t)(keep) * sizeof(type));      \
    free(ptr);                                                                          \
    (ptr) = fresh_;                                                                     \
  } while (0)

static int grow_ghosts(Env* s, int need) {
  if (need <= s->gcap) return 0;
  int cap = s->gcap ? s->gcap : 64;
  while (cap < need) cap *= 2;
  const int g = s->g;
  ETP_REGROW(s->gids, PyObject*, cap, g);   ETP_REGROW(s->gids2, PyObject*, cap, 0);
  ETP_REGROW(s->gten, PyObject*, cap, s->gten_n); ETP_REGROW(s->gten2, PyObject*, cap, 0);
  ETP_REGROW(s->flens, int32_t, cap, g);    ETP_REGROW(s->flens2, int32_t, cap, 0);
  ETP_REGROW(s->fcum, int32_t, cap, g);     ETP_REGROW(s->fcum2, int32_t, cap, 0);
  ETP_REGROW(s->gpos, double, 3 * cap, 3 * g); ETP_REGROW(s->gpos2, double, 3 * cap, 0);
  ETP_REGROW(s->gptr, int64_t, cap, g);     ETP_REGROW(s->gptr2, int64_t, cap, 0);
  ETP_REGROW(s->gwt, float, cap, g);        ETP_REGROW(s->gwt2, float, cap, 0);
  s->gcap = cap;
  return 0;
}
static int grow_fronts(Env* s, int need) {
  if (need <= s->nnzcap) return 0;
  int cap = s->nnzcap ? s->nnzcap : 128;
  while (cap < need) cap *= 2;
  ETP_REGROW(s->fidx, int32_t, cap, s->nnz); ETP_REGROW(s->fidx2, int32_t, cap, 0);
  s->nnzcap = cap;
  return 0;
}

static int as_double(PyObject* o, double* out) {
  if (PyFloat_CheckExact(o)) { *out = PyFloat_AS_DOUBLE(o); return 0; }
  double v = PyFloat_AsDouble(o);
  if (v == -1.0 && PyErr_Occurred()) {           /* e.g. a 0-d tensor: goes through __float__ */
    PyErr_Clear();
    PyObject* f = PyNumber_Float(o);
    if (!f) return -1;
    v = PyFloat_AS_DOUBLE(f);
    Py_DECREF(f);
  }
  *out = v;
  return 0;
}
/* three coordinates of a position (numpy float64 / float32 array through the buffer protocol, anything else as a sequence) */
static int read3(PyObject* o, double* out) {
  Py_buffer v;
  if (PyObject_CheckBuffer(o) && PyObject_GetBuffer(o, &v, PyBUF_FORMAT | PyBUF_STRIDES) == 0) {
    int ok = 0;
    if (v.ndim == 1 && v.shape[0] == 3 && v.format) {
      const char* f = v.format;
      if (f[0] == '<' || f[0] == '=' || f[0] == '@') ++f;
      if (f[0] == 'd' && f[1] == 0) { for (int k = 0; k < 3; ++k) out[k] = *(const double*)((const char*)v.buf + k * v.strides[0]); ok = 1; }
      else if (f[0] == 'f' && f[1] == 0) { for (int k = 0; k < 3; ++k) out[k] = (double)*(const float*)((const char*)v.buf + k * v.strides[0]); ok = 1; }
    }
    PyBuffer_Release(&v);
    if (ok) return 0;
  } else {
    PyErr_Clear();
  }
  for (int k = 0; k < 3; ++k) {
    PyObject* it = PySequence_GetItem(o, k);
    if (!it) return -1;
    int e = as_double(it, out + k);
    Py_DECREF(it);
    if (e) return -1;
  }
  return 0;
}
static int same_key(PyObject* a, PyObject* b) {
  if (a == b) return 1;
  int r = PyObject_RichCompareBool(a, b, Py_EQ);
  return r;   /* -1 on error */
}
static PyObject* getitem_dict(PyObject* d, PyObject* k) {   /* borrowed for dicts, NULL + KeyError when missing */
  PyObject* v = PyDict_GetItemWithError(d, k);
  if (!v && !PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, k);
  return v;
}
static Py_ssize_t len_of(PyObject* o) { return PyList_CheckExact(o) ? PyList_GET_SIZE(o) : PyObject_Size(o); }

/* tables[i][j0..j1) from the dictionaries (row dictionaries looked up once) */
static int read_row(Env* s, PyObject* sd, PyObject* sp, int i, int j0, int j1) {
  PyObject *ri = getitem_dict(sd, s->nid[i]), *pi = ri ? getitem_dict(sp, s->nid[i]) : NULL;
  if (!ri || !pi) return -1;
  if (!PyDict_Check(ri) || !PyDict_Check(pi)) { PyErr_SetString(PyExc_TypeError, "shortest_dist / shortest_path rows must be dicts"); return -1; }
  for (int j = j0; j < j1; ++j) {
    PyObject *d = getitem_dict(ri, s->nid[j]), *p = d ? getitem_dict(pi, s->nid[j]) : NULL;
    if (!d || !p) return -1;
    if (as_double(d, s->sd + (size_t)i * s->cap + j)) return -1;
    Py_ssize_t l = len_of(p);
    if (l < 0) return -1;
    s->spl[(size_t)i * s->cap + j] = (int32_t)l;
  }
  return 0;
}

/* Device pointer of an embedding row the image gather may read in place, or -1: ctx = (device, dtype, width, grad enabled).
 * The row must live on that device, be a contiguous 1-D tensor of `width` elements of that dtype, and — while autograd is
 * recording — not require a gradient (`usable` in packing._EnvMirror.sync is the same rule in Python). */
static long long row_pointer(PyObject* t, PyObject* ctx) {
  static PyObject *s_dtype, *s_device, *s_ndim, *s_shape, *s_contig, *s_rg, *s_ptr;
  if (!s_ptr) {
    s_dtype = PyUnicode_InternFromString("dtype"); s_device = PyUnicode_InternFromString("device");
    s_ndim = PyUnicode_InternFromString("ndim"); s_shape = PyUnicode_InternFromString("shape");
    s_contig = PyUnicode_InternFromString("is_contiguous"); s_rg = PyUnicode_InternFromString("requires_grad");
    s_ptr = PyUnicode_InternFromString("data_ptr");
  }
  long long out = -1;
  PyObject *a = NULL, *b = NULL;
  a = PyObject_GetAttr(t, s_dtype);
  if (!a || a != PyTuple_GET_ITEM(ctx, 1)) goto no;
  Py_CLEAR(a);
  a = PyObject_GetAttr(t, s_device);
  if (!a) goto no;
  { int eq = PyObject_RichCompareBool(a, PyTuple_GET_ITEM(ctx, 0), Py_EQ); if (eq != 1) goto no; }
  Py_CLEAR(a);
  a = PyObject_GetAttr(t, s_ndim);
  if (!a || PyLong_AsLong(a) != 1) goto no;
  Py_CLEAR(a);
  a = PyObject_GetAttr(t, s_shape);
  b = a ? PySequence_GetItem(a, 0) : NULL;
  if (!b || PyLong_AsLong(b) != PyLong_AsLong(PyTuple_GET_ITEM(ctx, 2))) goto no;
  Py_CLEAR(a); Py_CLEAR(b);
  a = PyObject_CallMethodNoArgs(t, s_contig);
  if (!a || a != Py_True) goto no;
  Py_CLEAR(a);
  if (PyTuple_GET_ITEM(ctx, 3) == Py_True) {
    a = PyObject_GetAttr(t, s_rg);
    if (!a || a != Py_False) goto no;
    Py_CLEAR(a);
  }
  a = PyObject_CallMethodNoArgs(t, s_ptr);
  if (a) out = PyLong_AsLongLong(a);
no:
  Py_XDECREF(a); Py_XDECREF(b);
  if (PyErr_Occurred()) { PyErr_Clear(); out = -1; }
  return out;
}

static int sync_env(Env* s, PyObject* gm, int want_img, PyObject* ptr_of) {
  int rc = -1, updated;
  PyObject *sd = NULL, *sp = NULL, *node_pos = NULL, *node_step = NULL, *gpos_d = NULL, *fronts_d = NULL, *aug_d = NULL,
           *G = NULL, *nemb = NULL, *gemb = NULL;
  sd = PyObject_GetAttrString(gm, "shortest_dist");
  if (!sd) goto done;
  if (!PyDict_Check(sd)) { PyErr_SetString(PyExc_TypeError, "GraphMap.shortest_dist is not a dict (pack before the first update_graph?)"); goto done; }
  updated = (sd != s->gm_sd);
  if (updated) {
    node_pos = PyObject_GetAttrString(gm, "node_pos");
    node_step = PyObject_GetAttrString(gm, "node_stepId");
    sp = PyObject_GetAttrString(gm, "shortest_path");
    if (!node_pos || !node_step || !sp || !PyDict_Check(node_pos) || !PyDict_Check(sp)) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_TypeError, "GraphMap tables must be dicts"); goto done; }
    int n = (int)PyDict_Size(node_pos), n0 = s->n;
    /* the history update_graph produces is append-only: otherwise start over */
    PyObject *key, *val; Py_ssize_t pos = 0; int k = 0, prefix_ok = (n >= n0);
    while (prefix_ok && k < n0 && PyDict_Next(node_pos, &pos, &key, &val)) {
      int e = same_key(key, s->nid[k]);
      if (e < 0) goto done;
      if (!e) prefix_ok = 0;
      ++k;
    }
    if (!prefix_ok) {
      env_clear(s);
      if (env_init(s)) goto done;
      n0 = 0;
    }
    if (grow_nodes(s, n)) goto done;
    pos = 0; k = 0;
    while (PyDict_Next(node_pos, &pos, &key, &val)) {
      if (k >= n0) {
        Py_INCREF(key); s->nid[k] = key;
        PyObject* idx = PyLong_FromLong(k);
        if (!idx || PyDict_SetItem(s->ix, key, idx)) { Py_XDECREF(idx); s->n = k + 1; goto done; }
        Py_DECREF(idx);
        s->n = k + 1;
        if (read3(val, s->npos + 3 * k)) goto done;
        PyObject* st = PyObject_GetItem(node_step, key);
        if (!st) goto done;
        long sv = PyLong_AsLong(st);
        if (sv == -1 && PyErr_Occurred()) { PyErr_Clear(); PyObject* l = PyNumber_Long(st); if (!l) { Py_DECREF(st); goto done; } sv = PyLong_AsLong(l); Py_DECREF(l); }
        Py_DECREF(st);
        s->steps[k] = (int32_t)sv;
      }
      ++k;
    }
    s->steps[n] = 0;
    int leaf = 0;
    if (n0 > 0 && n == n0 + 1 && PyObject_HasAttrString(gm, "graph_nx")) {
      G = PyObject_GetAttrString(gm, "graph_nx");
      /* degree of the new node: networkx keeps {node: {neighbour: attrs}} in Graph._adj (read directly: G[v] builds a view
       * object in Python for every call); any other graph type goes through G[v] */
      PyObject* adjd = G ? PyObject_GetAttrString(G, "_adj") : NULL;
      PyObject* nb = (adjd && PyDict_Check(adjd)) ? PyDict_GetItemWithError(adjd, s->nid[n0]) : NULL;   /* borrowed */
      if (nb && PyDict_Check(nb)) {
        leaf = (PyDict_Size(nb) == 1);
      } else {
        if (PyErr_Occurred()) PyErr_Clear();
        PyObject* adj = G ? PyObject_GetItem(G, s->nid[n0]) : NULL;
        if (adj) { leaf = (PyObject_Size(adj) == 1); Py_DECREF(adj); }
      }
      Py_XDECREF(adjd);
      if (PyErr_Occurred()) PyErr_Clear();
    }
    if (n0 > 0 && n == n0 + 1) {
      /* ONE new node v: its row and its column are new.  Between two OLDER nodes a, b networkx's Dijkstra (strict `<`
       * relaxation, _dijkstra_multisource) keeps distance, path and tie-breaks unless a route through v is not longer than
       * what the pair had: d(a,v) + d(v,b) <= old d(a,b) (with a 1e-9 relative margin for the rounding of the two sums).
       * Only those pairs are read again — none when v has degree 1 (no simple path passes through a leaf), a handful after a
       * loop closure.  ETP_PACK_VERIFY=1 (tests) re-reads everything afterwards and raises on any difference. */
      if (read_row(s, sd, sp, n0, 0, n)) goto done;
      for (int i = 0; i < n0; ++i) if (read_row(s, sd, sp, i, n0, n)) goto done;
      if (!leaf) {
        for (int i = 0; i < n0; ++i) {
          const double av = s->sd[(size_t)i * s->cap + n0];
          for (int j = 0; j < n0; ++j) {
            if (i == j) continue;
            const double cand = av + s->sd[(size_t)n0 * s->cap + j], was = s->sd[(size_t)i * s->cap + j];
            if (!(cand > was * (1.0 + 1e-9)) && read_row(s, sd, sp, i, j, j + 1)) goto done;
          }
        }
      }
      const char* vf = getenv("ETP_PACK_VERIFY");
      if (vf && vf[0] == '1') {
        double* keep_d = (double*)malloc((size_t)n * n * 8);
        int32_t* keep_l = (int32_t*)malloc((size_t)n * n * 4);
        if (!keep_d || !keep_l) { free(keep_d); free(keep_l); PyErr_NoMemory(); goto done; }
        for (int i = 0; i < n; ++i) { memcpy(keep_d + (size_t)i * n, s->sd + (size_t)i * s->cap, (size_t)n * 8); memcpy(keep_l + (size_t)i * n, s->spl + (size_t)i * s->cap, (size_t)n * 4); }
        int bad = 0;
        for (int i = 0; i < n && !bad; ++i) if (read_row(s, sd, sp, i, 0, n)) bad = -1;
        for (int i = 0; i < n && !bad; ++i)
          if (memcmp(keep_d + (size_t)i * n, s->sd + (size_t)i * s->cap, (size_t)n * 8) || memcmp(keep_l + (size_t)i * n, s->spl + (size_t)i * s->cap, (size_t)n * 4)) bad = 1;
        free(keep_d); free(keep_l);
        if (bad < 0) goto done;
        if (bad) { PyErr_SetString(PyExc_RuntimeError, "gmap mirror: the selective table update differs from a full read"); goto done; }
      }
    } else {
      for (int i = 0; i < n; ++i) if (read_row(s, sd, sp, i, 0, n)) goto done;
    }
    Py_INCREF(sd);
    Py_XSETREF(s->gm_sd, sd);
  }
  /* ghosts */
  gpos_d = PyObject_GetAttrString(gm, "ghost_pos");
  fronts_d = PyObject_GetAttrString(gm, "ghost_fronts");
  if (!gpos_d || !fronts_d || !PyDict_Check(gpos_d) || !PyDict_Check(fronts_d)) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_TypeError, "ghost tables must be dicts"); goto done; }
  {
    const int g = (int)PyDict_Size(gpos_d), g0 = s->g;
    if (grow_ghosts(s, g)) goto done;
    PyObject **ng = s->gids2, **og = s->gids;
    int32_t *nl = s->flens2, *ol = s->flens;
    int* from = (int*)malloc(sizeof(int) * (g ? g : 1));     /* index of ghost i in the previous state, or -1 */
    if (!from) { PyErr_NoMemory(); goto done; }
    PyObject *key, *val; Py_ssize_t pos = 0; int k = 0, j = 0, same_ids = (g == g0), changed, nnz = 0;
    /* both sequences are in creation order (ids are never reused; deletions only remove): one forward walk finds every
     * surviving ghost; ids are compared by identity (a different-but-equal string object just re-reads that ghost) */
    while (PyDict_Next(gpos_d, &pos, &key, &val)) {
      int jj = j;
      while (jj < g0 && og[jj] != key) ++jj;
      if (jj < g0) { from[k] = jj; j = jj + 1; } else from[k] = -1;
      if (from[k] != k) same_ids = 0;
      ng[k] = key;
      PyObject* fl = getitem_dict(fronts_d, key);
      Py_ssize_t l = fl ? len_of(fl) : -1;
      if (l < 0) { free(from); goto done; }
      nl[k] = (int32_t)l; nnz += (int)l;
      ++k;
    }
    changed = !same_ids;
    if (same_ids) for (int i = 0; i < g; ++i) if (nl[i] != ol[i]) { changed = 1; break; }
    if (changed) {      /* front CSR: segments of unchanged ghosts are copied, the others re-read */
      if (grow_fronts(s, nnz)) { free(from); goto done; }
      int w = 0;
      for (int i = 0; i < g; ++i) {
        const int f = from[i];
        if (f >= 0 && ol[f] == nl[i]) {
          const int o0 = f ? s->fcum[f - 1] : 0;
          memcpy(s->fidx2 + w, s->fidx + o0, (size_t)nl[i] * 4);
          w += nl[i];
        } else {
          PyObject* fl = getitem_dict(fronts_d, ng[i]);
          PyObject* fast = fl ? PySequence_Fast(fl, "ghost_fronts entries must be sequences") : NULL;
          if (!fast) { free(from); goto done; }
          for (Py_ssize_t q = 0; q < PySequence_Fast_GET_SIZE(fast); ++q) {
            PyObject* ixv = getitem_dict(s->ix, PySequence_Fast_GET_ITEM(fast, q));
            if (!ixv) { Py_DECREF(fast); free(from); goto done; }
            s->fidx2[w++] = (int32_t)PyLong_AsLong(ixv);
          }
          Py_DECREF(fast);
        }
        s->fcum2[i] = w;
      }
    }
    const int repos = updated || !same_ids;
    if (repos) {
      /* ghost_aug_pos is redrawn for EVERY ghost by update_graph when ghost_aug != 0 (:238-244); with ghost_aug == 0 it is
       * the mean position, which moves only when the ghost gains a front; without an update nothing moves */
      int frozen = !updated;
      if (updated && PyObject_HasAttrString(gm, "ghost_aug")) {
        PyObject* ga = PyObject_GetAttrString(gm, "ghost_aug");
        double gv = 1.0;
        if (ga && as_double(ga, &gv) == 0 && gv == 0.0) frozen = 1;
        Py_XDECREF(ga);
        if (PyErr_Occurred()) PyErr_Clear();
      }
      for (int i = 0; i < g; ++i) {
        const int f = from[i];
        if (f >= 0 && frozen && (!updated || ol[f] == nl[i])) { memcpy(s->gpos2 + 3 * i, s->gpos + 3 * f, 24); continue; }
        if (!aug_d) { aug_d = PyObject_GetAttrString(gm, "ghost_aug_pos"); if (!aug_d) { free(from); goto done; } }
        PyObject* p = PyObject_GetItem(aug_d, ng[i]);
        if (!p || read3(p, s->gpos2 + 3 * i)) { Py_XDECREF(p); free(from); goto done; }
        Py_DECREF(p);
      }
    }
    /* image rows: pointers / weights of the per-node and per-ghost embedding tensors */
    int reimg = 0;
    if (!want_img && changed) s->img_g = -1;     /* ghost rows of an earlier image sync are stale now */
    if (want_img && s->img_ok) {
      if (s->nimg < s->n) {
        nemb = PyObject_GetAttrString(gm, "node_embeds");
        if (!nemb) { free(from); goto done; }
        for (int q = s->nimg; q < s->n && s->img_ok; ++q) {
          PyObject* t = PyObject_GetItem(nemb, s->nid[q]);
          if (!t) { free(from); goto done; }
          long long pv = row_pointer(t, ptr_of);
          if (pv < 0) { s->img_ok = 0; Py_DECREF(t); break; }
          s->nptr[q] = pv;
          PyList_Append(s->nkeep, t);
          Py_DECREF(t);
          s->nimg = q + 1;
        }
      }
      if (s->img_ok && (changed || s->img_g != g0)) {
        const int old_valid = (s->img_g == g0);       /* the previous state's ghost rows are usable */
        reimg = 1;
        for (int i = 0; i < g && s->img_ok; ++i) {
          const int f = from[i];
          if (old_valid && f >= 0 && ol[f] == nl[i]) {  /* ghost_embeds[g] is replaced exactly when ghost_fronts[g] grows */
            s->gptr2[i] = s->gptr[f]; s->gwt2[i] = s->gwt[f];
            s->gten2[i] = s->gten[f]; Py_XINCREF(s->gten2[i]);
            continue;
          }
          if (!gemb) { gemb = PyObject_GetAttrString(gm, "ghost_embeds"); if (!gemb) { for (int q = 0; q < i; ++q) Py_CLEAR(s->gten2[q]); free(from); goto done; } }
          PyObject* ent = PyObject_GetItem(gemb, ng[i]);            /* [sum tensor, count] */
          PyObject* t = ent ? PySequence_GetItem(ent, 0) : NULL;
          PyObject* cnt = ent ? PySequence_GetItem(ent, 1) : NULL;
          double cv = 0.0;
          if (!t || !cnt || as_double(cnt, &cv)) {
            Py_XDECREF(ent); Py_XDECREF(t); Py_XDECREF(cnt);
            for (int q = 0; q < i; ++q) Py_CLEAR(s->gten2[q]);
            free(from); goto done;
          }
          long long pv = row_pointer(t, ptr_of);
          Py_DECREF(ent); Py_DECREF(cnt);
          if (pv < 0) { s->img_ok = 0; Py_DECREF(t); s->gten2[i] = NULL; for (int q = 0; q < i; ++q) Py_CLEAR(s->gten2[q]); reimg = 0; break; }
          s->gptr2[i] = pv;
          s->gwt2[i] = (float)(1.0 / cv);
          s->gten2[i] = t;                                          /* keeps the row alive while its pointer is in the table */
        }
      }
    }
    /* swap the new state in */
    for (int i = 0; i < g; ++i) Py_INCREF(ng[i]);
    for (int i = 0; i < g0; ++i) Py_DECREF(og[i]);
    s->gids = ng; s->gids2 = og; s->flens = nl; s->flens2 = ol;
    if (changed) { int32_t* t1 = s->fcum; s->fcum = s->fcum2; s->fcum2 = t1; t1 = s->fidx; s->fidx = s->fidx2; s->fidx2 = t1; s->nnz = nnz; }
    if (repos) { double* t2 = s->gpos; s->gpos = s->gpos2; s->gpos2 = t2; }
    if (reimg) {
      for (int i = 0; i < s->gten_n; ++i) Py_CLEAR(s->gten[i]);
      s->gten_n = g;
      int64_t* t3 = s->gptr; s->gptr = s->gptr2; s->gptr2 = t3;
      float* t4 = s->gwt; s->gwt = s->gwt2; s->gwt2 = t4;
      PyObject** t5 = s->gten; s->gten = s->gten2; s->gten2 = t5;
      s->img_g = g;
    } else if (!same_ids && s->img_g >= 0 && !(want_img && s->img_ok)) {
      s->img_g = -1;
    }
    s->g = g;
    if (g == 0) s->nnz = 0;
    free(from);
  }
  rc = 0;
done:
  Py_XDECREF(sd); Py_XDECREF(sp); Py_XDECREF(node_pos); Py_XDECREF(node_step); Py_XDECREF(gpos_d); Py_XDECREF(fronts_d);
  Py_XDECREF(aug_d); Py_XDECREF(G); Py_XDECREF(nemb); Py_XDECREF(gemb);
  return rc;
}

static Env* env_of(PyObject* cap) { return (Env*)PyCapsule_GetPointer(cap, "etp.gmap_mirror"); }

PyObject* etp_pm_new(void) {
  Env* s = (Env*)malloc(sizeof(Env));
  if (!s) return PyErr_NoMemory();
  if (env_init(s)) { free(s); return NULL; }
  s->img_g = -1;
  return PyCapsule_New(s, "etp.gmap_mirror", env_destroy);
}

/* sync every map; ctx = (device, dtype, width, grad enabled) of the rows the image gather may read in place;
 * returns (n_max, max_g, n_f64, n_i32, image rows, all maps image-fast) */
PyObject* etp_pm_sync(PyObject* states, PyObject* gmaps, int want_img, PyObject* ptr_of) {
  if (!PyTuple_Check(ptr_of) || PyTuple_GET_SIZE(ptr_of) != 4) { PyErr_SetString(PyExc_TypeError, "ctx must be (device, dtype, width, grad_enabled)"); return NULL; }
  Py_ssize_t B = PyList_Size(states);
  if (B < 0 || PyList_Size(gmaps) != B) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "states / gmaps length mismatch"); return NULL; }
  long n_max = 1, max_g = 0, nf = 0, ni = 0, rows = 0; int ok = 1;
  for (Py_ssize_t e = 0; e < B; ++e) {
    Env* s = env_of(PyList_GET_ITEM(states, e));
    if (!s || sync_env(s, PyList_GET_ITEM(gmaps, e), want_img, ptr_of)) return NULL;
    long n = s->n, g = s->g;
    if (1 + n + g > n_max) n_max = 1 + n + g;
    if (g > max_g) max_g = g;
    nf += 4 + 3 * n + 3 * g + n * n;
    ni += n + 1 + g + s->nnz + n * n;
    rows += n + g;
    ok = ok && s->img_ok;
  }
  return Py_BuildValue("(lllllO)", n_max, max_g, nf, ni, rows, ok ? Py_True : Py_False);
}

/* write meta / the two blobs / (optionally) the row-pointer table and the CSR of the image gather at the given addresses */
PyObject* etp_pm_fill(PyObject* states, PyObject* cur_vp, const double* poses, int32_t* meta, double* fout, int32_t* iout,
                      int64_t* table, int32_t* cptr, int32_t* cidx, float* cwt, int n_max) {
  Py_ssize_t B = PyList_Size(states);
  if (B < 0 || PyList_Size(cur_vp) != B) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "states / cur_vp length mismatch"); return NULL; }
  long of = 0, oi = 0, row = 0;
  for (Py_ssize_t e = 0; e < B; ++e) {
    Env* s = env_of(PyList_GET_ITEM(states, e));
    if (!s) return NULL;
    const int n = s->n, g = s->g;
    PyObject* ci = getitem_dict(s->ix, PyList_GET_ITEM(cur_vp, e));
    if (!ci) return NULL;
    int32_t* m = meta + 8 * e;
    m[0] = n; m[1] = g; m[2] = (int32_t)PyLong_AsLong(ci); m[3] = (int32_t)of; m[4] = (int32_t)oi; m[5] = s->nnz; m[6] = 0; m[7] = 0;
    double* f = fout + of;
    memcpy(f, poses + 4 * e, 32); f += 4;
    memcpy(f, s->npos, (size_t)n * 24); f += 3 * n;
    if (g) memcpy(f, s->gpos, (size_t)g * 24);
    f += 3 * g;
    for (int i = 0; i < n; ++i) { memcpy(f, s->sd + (size_t)i * s->cap, (size_t)n * 8); f += n; }
    of += 4 + 3 * n + 3 * g + (long)n * n;
    int32_t* q = iout + oi;
    memcpy(q, s->steps, (size_t)(n + 1) * 4); q += n + 1;
    if (g) memcpy(q, s->fcum, (size_t)g * 4);
    q += g;
    if (s->nnz) memcpy(q, s->fidx, (size_t)s->nnz * 4);
    q += s->nnz;
    for (int i = 0; i < n; ++i) { memcpy(q, s->spl + (size_t)i * s->cap, (size_t)n * 4); q += n; }
    oi += n + 1 + g + s->nnz + (long)n * n;
    if (table) {
      const int k = n + g;
      memcpy(table + row, s->nptr, (size_t)n * 8);
      if (g) memcpy(table + row + n, s->gptr, (size_t)g * 8);
      for (int j = 0; j < n; ++j) cwt[row + j] = 1.0f;
      if (g) memcpy(cwt + row + n, s->gwt, (size_t)g * 4);
      for (int j = 0; j < k; ++j) cidx[row + j] = (int32_t)(row + j);
      int32_t* p = cptr + (size_t)e * n_max;
      for (int j = 0; j < n_max; ++j) { int c = j - 1; if (c < 0) c = 0; if (c > k) c = k; p[j] = (int32_t)(row + c); }
      row += k;
    }
  }
  if (table) cptr[(size_t)B * n_max] = (int32_t)row;
  Py_RETURN_NONE;
}

/* ([None] + node ids + ghost ids per map, [no ghosts left per map]) */
PyObject* etp_pm_vp_ids(PyObject* states) {
  Py_ssize_t B = PyList_Size(states);
  if (B < 0) return NULL;
  PyObject *all = PyList_New(B), *none_left = PyList_New(B);
  if (!all || !none_left) { Py_XDECREF(all); Py_XDECREF(none_left); return NULL; }
  for (Py_ssize_t e = 0; e < B; ++e) {
    Env* s = env_of(PyList_GET_ITEM(states, e));
    if (!s) { Py_DECREF(all); Py_DECREF(none_left); return NULL; }
    PyObject* l = PyList_New(1 + s->n + s->g);
    if (!l) { Py_DECREF(all); Py_DECREF(none_left); return NULL; }
    Py_INCREF(Py_None); PyList_SET_ITEM(l, 0, Py_None);
    for (int i = 0; i < s->n; ++i) { Py_INCREF(s->nid[i]); PyList_SET_ITEM(l, 1 + i, s->nid[i]); }
    for (int i = 0; i < s->g; ++i) { Py_INCREF(s->gids[i]); PyList_SET_ITEM(l, 1 + s->n + i, s->gids[i]); }
    PyList_SET_ITEM(all, e, l);
    PyObject* b = s->g == 0 ? Py_True : Py_False;
    Py_INCREF(b); PyList_SET_ITEM(none_left, e, b);
  }
  return Py_BuildValue("(NN)", all, none_left);
}
