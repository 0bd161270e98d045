/*
 * pf_text.c — VPR's .route / .place text files from the flat data model (pf_text.h).
 *
 * Writers reproduce the reference's fprintf sequences byte for byte (print_route route_common.c:1322-1417,
 * print_place read_place.c:266-293) but format into a large buffer with a hand-rolled integer printer:
 * a 200 k-net routing is ~10^7 "Node:" lines and fprintf would dominate the turn-around.  Readers parse a
 * whole file from memory; pf_place_read finds blocks through a hash table where read_place.c:108-114 scans
 * the block list with strcmp for every line.
 * Plain C host code; no CUDA.
 */
#include "pf_text.h"

#include <ctype.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static const char NAME_MAGIC[8] = { 'P', 'F', 'N', 'A', 'M', 'E', '0', '1' };

static __thread char g_err[256];
const char *pf_text_error(void) { return g_err; }
static int fail(int code, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

/* ------------------------------------------------------------------ pf_names container */
static int wr(FILE *f, const void *p, size_t bytes) {
	if (bytes == 0) return 0;
	return fwrite(p, 1, bytes, f) == bytes ? 0 : PF_EIO;
}
static int rd_alloc(FILE *f, void **pp, size_t bytes) {
	*pp = NULL;
	if (bytes == 0) return 0;
	*pp = malloc(bytes);
	if (!*pp) return PF_ENOMEM;
	return fread(*pp, 1, bytes, f) == bytes ? 0 : PF_EIO;
}
#define W(ptr, n) do { if ((rc = wr(f, (ptr), (size_t)(n) * sizeof(*(ptr)))) != 0) goto done; } while (0)
#define R(ptr, n) do { if ((rc = rd_alloc(f, (void **)&(ptr), (size_t)(n) * sizeof(*(ptr)))) != 0) goto done; } while (0)

static size_t tiles_of(const pf_names *n) { return (size_t)(n->nx + 2) * (size_t)(n->ny + 2); }

int pf_names_write(const char *path, const pf_names *n) {
	int rc = 0;
	int32_t hdr[16];
	FILE *f = fopen(path, "wb");
	if (!f) return PF_EIO;
	memset(hdr, 0, sizeof(hdr));
	hdr[0] = n->nx; hdr[1] = n->ny; hdr[2] = n->num_nets; hdr[3] = n->num_blocks;
	hdr[4] = n->net_name_ptr[n->num_nets];
	hdr[5] = n->num_blocks ? n->block_name_ptr[n->num_blocks] : 0;
	hdr[6] = n->gpin_ptr[n->num_nets];
	if ((rc = wr(f, NAME_MAGIC, 8)) != 0) goto done;
	if ((rc = wr(f, hdr, sizeof(hdr))) != 0) goto done;
	W(n->net_name_ptr, (size_t)n->num_nets + 1);
	W(n->net_name_chars, hdr[4]);
	W(n->tile_is_io, tiles_of(n));
	W(n->block_name_ptr, (size_t)n->num_blocks + 1);
	W(n->block_name_chars, hdr[5]);
	W(n->block_x, n->num_blocks); W(n->block_y, n->num_blocks); W(n->block_z, n->num_blocks);
	W(n->gpin_ptr, (size_t)n->num_nets + 1);
	W(n->gpin_block, hdr[6]); W(n->gpin_class, hdr[6]);
done:
	if (fclose(f) != 0 && rc == 0) rc = PF_EIO;
	return rc;
}

int pf_names_read(const char *path, pf_names *n) {
	int rc = 0;
	int32_t hdr[16];
	char magic[8];
	FILE *f = fopen(path, "rb");
	memset(n, 0, sizeof(*n));
	if (!f) return PF_EIO;
	if (fread(magic, 1, 8, f) != 8 || memcmp(magic, NAME_MAGIC, 8) != 0) { rc = PF_EFORMAT; goto done; }
	if (fread(hdr, 1, sizeof(hdr), f) != sizeof(hdr)) { rc = PF_EIO; goto done; }
	n->nx = hdr[0]; n->ny = hdr[1]; n->num_nets = hdr[2]; n->num_blocks = hdr[3];
	if (n->nx < 0 || n->ny < 0 || n->num_nets < 0 || n->num_blocks < 0 || hdr[4] < 0 || hdr[5] < 0 || hdr[6] < 0) {
		rc = PF_EFORMAT; goto done;
	}
	R(n->net_name_ptr, (size_t)n->num_nets + 1);
	R(n->net_name_chars, hdr[4]);
	R(n->tile_is_io, tiles_of(n));
	R(n->block_name_ptr, (size_t)n->num_blocks + 1);
	R(n->block_name_chars, hdr[5]);
	R(n->block_x, n->num_blocks); R(n->block_y, n->num_blocks); R(n->block_z, n->num_blocks);
	R(n->gpin_ptr, (size_t)n->num_nets + 1);
	R(n->gpin_block, hdr[6]); R(n->gpin_class, hdr[6]);
	if (n->net_name_ptr[n->num_nets] != hdr[4] || n->block_name_ptr[n->num_blocks] != hdr[5]
			|| n->gpin_ptr[n->num_nets] != hdr[6]) rc = PF_EFORMAT;
done:
	fclose(f);
	if (rc != 0) pf_names_free(n);
	return rc;
}

void pf_names_free(pf_names *n) {
	free(n->net_name_ptr); free(n->net_name_chars); free(n->tile_is_io);
	free(n->block_name_ptr); free(n->block_name_chars);
	free(n->block_x); free(n->block_y); free(n->block_z);
	free(n->gpin_ptr); free(n->gpin_block); free(n->gpin_class);
	memset(n, 0, sizeof(*n));
}

#define FAIL(...) do { if (msg && msg_len > 0) snprintf(msg, (size_t)msg_len, __VA_ARGS__); return PF_EINVAL; } while (0)

static int names_ok(const int32_t *ptr, const char *chars, int count, const char *what, char *msg, int msg_len) {
	int i, k;
	if (ptr[0] != 0) FAIL("%s name offsets do not start at 0", what);
	for (i = 0; i < count; i++)                      /* all offsets first: the character array is ptr[count] long */
		if (ptr[i + 1] <= ptr[i]) FAIL("%s %d has an empty name", what, i);
	for (i = 0; i < count; i++) {
		for (k = ptr[i]; k < ptr[i + 1]; k++)
			if (chars[k] == 0 || chars[k] == '\n' || chars[k] == '\r' || chars[k] == ' ' || chars[k] == '\t')
				FAIL("%s %d: white space or NUL in the name", what, i);
	}
	return PF_OK;
}

int pf_names_check(const pf_names *n, const pf_problem *p, char *msg, int msg_len) {
	int i, k, rc;
	if (msg && msg_len > 0) msg[0] = 0;
	if (p && (n->nx != p->nx || n->ny != p->ny)) FAIL("grid %d x %d, problem has %d x %d", n->nx, n->ny, p->nx, p->ny);
	if (p && n->num_nets != p->num_nets) FAIL("%d net names for %d nets", n->num_nets, p->num_nets);
	if ((rc = names_ok(n->net_name_ptr, n->net_name_chars, n->num_nets, "net", msg, msg_len)) != 0) return rc;
	if (n->num_blocks > 0 && (rc = names_ok(n->block_name_ptr, n->block_name_chars, n->num_blocks, "block", msg, msg_len)) != 0)
		return rc;
	for (i = 0; i < n->num_blocks; i++)
		if (n->block_x[i] < 0 || n->block_x[i] > n->nx + 1 || n->block_y[i] < 0 || n->block_y[i] > n->ny + 1)
			FAIL("block %d at (%d,%d) outside the grid", i, n->block_x[i], n->block_y[i]);
	if (n->gpin_ptr[0] != 0) FAIL("gpin_ptr does not start at 0");
	for (i = 0; i < n->num_nets; i++)
		if (n->gpin_ptr[i + 1] < n->gpin_ptr[i]) FAIL("gpin_ptr not monotone at net %d", i);
	for (i = 0; i < n->num_nets; i++) {
		if (p && !p->net_is_global[i] && n->gpin_ptr[i + 1] != n->gpin_ptr[i]) FAIL("routed net %d lists global pins", i);
		for (k = n->gpin_ptr[i]; k < n->gpin_ptr[i + 1]; k++)
			if (n->gpin_block[k] < 0 || n->gpin_block[k] >= n->num_blocks) FAIL("net %d: pin block %d out of range", i, n->gpin_block[k]);
	}
	return PF_OK;
}

static int digits_of(int v) { int d = 1; while (v >= 10) { v /= 10; d++; } return d; }

int pf_names_synthetic(const pf_problem *p, pf_names *n) {
	int i, x, y;
	size_t chars = 0, T;
	memset(n, 0, sizeof(*n));
	n->nx = p->nx; n->ny = p->ny; n->num_nets = p->num_nets; n->num_blocks = 0;
	T = tiles_of(n);
	for (i = 0; i < p->num_nets; i++) chars += 1 + (size_t)digits_of(i);
	n->net_name_ptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)p->num_nets + 1));
	n->net_name_chars = (char *)malloc(chars + 16);
	n->tile_is_io = (uint8_t *)calloc(T, 1);
	n->block_name_ptr = (int32_t *)calloc(1, sizeof(int32_t));
	n->gpin_ptr = (int32_t *)calloc((size_t)p->num_nets + 1, sizeof(int32_t));
	if (!n->net_name_ptr || !n->net_name_chars || !n->tile_is_io || !n->block_name_ptr || !n->gpin_ptr) {
		pf_names_free(n);
		return PF_ENOMEM;
	}
	chars = 0;
	for (i = 0; i < p->num_nets; i++) {
		n->net_name_ptr[i] = (int32_t)chars;
		chars += (size_t)sprintf(n->net_name_chars + chars, "n%d", i);
	}
	n->net_name_ptr[p->num_nets] = (int32_t)chars;
	for (x = 0; x <= p->nx + 1; x++)
		for (y = 0; y <= p->ny + 1; y++)
			if (x == 0 || y == 0 || x == p->nx + 1 || y == p->ny + 1) n->tile_is_io[(size_t)x * (size_t)(p->ny + 2) + (size_t)y] = 1;
	return PF_OK;
}

/* ------------------------------------------------------------------ host threads
 * The text of a 200 k-net routing is ~4 x 10^6 "Node:" lines, each needing six random reads of the 18 M-entry node
 * arrays: one thread is bound by cache misses (1.6 s for the 224 MB file), so nets are formatted / parsed in chunks by
 * all host threads (PF_TEXT_THREADS overrides the count). */
typedef void (*par_fn)(void *ctx, int chunk);
typedef struct { par_fn fn; void *ctx; int nchunks; int next; } par_job;
static void *par_worker(void *arg) {
	par_job *j = (par_job *)arg;
	for (;;) {
		int k = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
		if (k >= j->nchunks) return NULL;
		j->fn(j->ctx, k);
	}
}
static int par_threads(void) {
	const char *e = getenv("PF_TEXT_THREADS");
	long t = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
	if (t < 1) t = 1;
	if (t > 64) t = 64;
	return (int)t;
}
static void par_for(int nchunks, par_fn fn, void *ctx) {
	pthread_t th[64];
	par_job j;
	int T = par_threads(), i, started = 0;
	j.fn = fn; j.ctx = ctx; j.nchunks = nchunks; j.next = 0;
	if (T > nchunks) T = nchunks;
	for (i = 0; i < T - 1; i++) { if (pthread_create(&th[started], NULL, par_worker, &j) != 0) break; started++; }
	par_worker(&j);                                  /* the caller works too (and alone, if no thread could start) */
	for (i = 0; i < started; i++) pthread_join(th[i], NULL);
}
/* first error of a parallel phase: workers have their own thread-local g_err */
typedef struct { int rc; char msg[256]; } par_err;
static void par_fail(par_err *e, int rc) {
	int expect = 0;
	if (__atomic_compare_exchange_n(&e->rc, &expect, rc, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) memcpy(e->msg, g_err, sizeof(e->msg));
}

/* ------------------------------------------------------------------ buffered text output
 * f != NULL: a 4 MB window flushed to the file; f == NULL: a growing memory buffer (one per chunk of nets) */
typedef struct {
	FILE *f;
	char *buf;
	size_t len, cap;
	int err;
} outbuf;

static int ob_open(outbuf *o, const char *path) {
	o->f = fopen(path, "wb");
	if (!o->f) return PF_EIO;
	o->cap = (size_t)4 << 20; o->len = 0; o->err = 0;
	o->buf = (char *)malloc(o->cap);
	if (!o->buf) { fclose(o->f); return PF_ENOMEM; }
	return PF_OK;
}
static int ob_mem_open(outbuf *o, size_t cap) {
	o->f = NULL; o->len = 0; o->err = 0; o->cap = cap < 4096 ? 4096 : cap;
	o->buf = (char *)malloc(o->cap);
	return o->buf ? PF_OK : PF_ENOMEM;
}
static void ob_flush(outbuf *o) {
	if (o->f && o->len && fwrite(o->buf, 1, o->len, o->f) != o->len) o->err = 1;
	if (o->f) o->len = 0;
}
static void ob_room_slow(outbuf *o, size_t need) {
	if (o->f) { ob_flush(o); return; }
	{
		size_t nc = o->cap * 2 > o->len + need ? o->cap * 2 : o->len + need + o->cap;
		char *nb = (char *)realloc(o->buf, nc);
		if (!nb) { o->err = 2; o->len = 0; return; }      /* keeps the buffer valid; the chunk is reported as PF_ENOMEM */
		o->buf = nb; o->cap = nc;
	}
}
static inline void ob_room(outbuf *o, size_t need) { if (o->len + need > o->cap) ob_room_slow(o, need); }
static inline void ob_mem(outbuf *o, const char *s, size_t n) {
	if (o->f && n > o->cap) { ob_flush(o); if (fwrite(s, 1, n, o->f) != n) o->err = 1; return; }
	ob_room(o, n);
	if (o->len + n > o->cap) return;                     /* failed growth */
	memcpy(o->buf + o->len, s, n); o->len += n;
}
#define ob_lit(o, s) ob_mem((o), (s), sizeof(s) - 1)
static inline void ob_int(outbuf *o, long v) {   /* "%d" */
	char t[24];
	int k = 0;
	unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
	ob_room(o, 24);
	if (o->len + 24 > o->cap) return;
	do { t[k++] = (char)('0' + u % 10); u /= 10; } while (u);
	if (v < 0) o->buf[o->len++] = '-';
	while (k) o->buf[o->len++] = t[--k];
}
static int ob_close(outbuf *o) {
	int rc;
	ob_flush(o);
	rc = o->err ? PF_EIO : PF_OK;
	if (fclose(o->f) != 0) rc = PF_EIO;
	free(o->buf);
	return rc;
}

/* ------------------------------------------------------------------ .route */
/* "%6s" of name_type[] (route_common.c:1329) */
static const char *const TYPE_PADDED[6] = { "SOURCE", "  SINK", "  IPIN", "  OPIN", " CHANX", " CHANY" };
static const char *const TYPE_NAME[6] = { "SOURCE", "SINK", "IPIN", "OPIN", "CHANX", "CHANY" };

static inline int tile_io(const pf_names *n, int x, int y) {
	if (x < 0 || y < 0 || x > n->nx + 1 || y > n->ny + 1) return 0;
	return n->tile_is_io[(size_t)x * (size_t)(n->ny + 2) + (size_t)y];
}

/* nets [lo, hi) in print_route's format (route_common.c:1333-1414) */
static int format_nets(outbuf *o, const pf_problem *p, const pf_names *n, const pf_result *r, int lo, int hi) {
	int inet, k;
	for (inet = lo; inet < hi; inet++) {
		const char *name = n->net_name_chars + n->net_name_ptr[inet];
		size_t name_len = (size_t)(n->net_name_ptr[inet + 1] - n->net_name_ptr[inet]);
		ob_lit(o, "\n\nNet "); ob_int(o, inet); ob_lit(o, " ("); ob_mem(o, name, name_len);
		if (p->net_is_global[inet]) {                                        /* :1394-1412 */
			ob_lit(o, "): global net connecting:\n\n");
			for (k = n->gpin_ptr[inet]; k < n->gpin_ptr[inet + 1]; k++) {
				int b = n->gpin_block[k];
				ob_lit(o, "Block ");
				ob_mem(o, n->block_name_chars + n->block_name_ptr[b], (size_t)(n->block_name_ptr[b + 1] - n->block_name_ptr[b]));
				ob_lit(o, " (#"); ob_int(o, b); ob_lit(o, ") at ("); ob_int(o, n->block_x[b]); ob_lit(o, ", ");
				ob_int(o, n->block_y[b]); ob_lit(o, "), Pin class "); ob_int(o, n->gpin_class[k]); ob_lit(o, ".\n");
			}
			continue;
		}
		ob_lit(o, ")\n\n");
		if (p->net_ptr[inet + 1] - p->net_ptr[inet] - 1 == 0) {               /* :1337-1339 */
			ob_lit(o, "\n\nUsed in local cluster only, reserved one CLB pin\n\n");
			continue;
		}
		for (k = r->trace_ptr[inet]; k < r->trace_ptr[inet + 1]; k++) {      /* :1344-1389 */
			int inode = r->trace_node[k], t, ilow, jlow;
			if (inode < 0 || inode >= p->num_nodes) return fail(PF_EINVAL, "net %d: trace node %d out of range", inet, inode);
			t = p->type[inode]; ilow = p->xlow[inode]; jlow = p->ylow[inode];
			if (t > PF_CHANY) return fail(PF_EINVAL, "net %d: unexpected traceback element type %d", inet, t);
			ob_lit(o, "Node:\t"); ob_int(o, inode); ob_lit(o, "\t"); ob_mem(o, TYPE_PADDED[t], 6);
			ob_lit(o, " ("); ob_int(o, ilow); ob_lit(o, ","); ob_int(o, jlow); ob_lit(o, ") ");
			if (ilow != p->xhigh[inode] || jlow != p->yhigh[inode]) {
				ob_lit(o, "to ("); ob_int(o, p->xhigh[inode]); ob_lit(o, ","); ob_int(o, p->yhigh[inode]); ob_lit(o, ") ");
			}
			if (t == PF_CHANX || t == PF_CHANY) ob_lit(o, " Track: ");
			else if (tile_io(n, ilow, jlow)) ob_lit(o, " Pad: ");
			else if (t == PF_IPIN || t == PF_OPIN) ob_lit(o, " Pin: ");
			else ob_lit(o, " Class: ");
			ob_int(o, p->ptc_num[inode]); ob_lit(o, "  \n");
		}
	}
	if (o->err == 2) return fail(PF_ENOMEM, "out of memory formatting nets %d..%d", lo, hi);
	return PF_OK;
}

/* net ranges of roughly equal trace length: bounds[0..nchunks], returns nchunks */
static int chunk_nets(const int32_t *trace_ptr, int num_nets, long per_chunk, int **bounds_out) {
	int cap = 16, nc = 0, inet, *b = (int *)malloc(sizeof(int) * (size_t)cap);
	long start_elems = 0;
	if (!b) return -1;
	b[0] = 0;
	for (inet = 0; inet < num_nets; inet++) {
		if ((long)trace_ptr[inet + 1] - start_elems + 8L * (inet + 1 - b[nc]) >= per_chunk || inet == num_nets - 1) {
			if (nc + 2 > cap) { int *nb = (int *)realloc(b, sizeof(int) * (size_t)(cap *= 2)); if (!nb) { free(b); return -1; } b = nb; }
			b[++nc] = inet + 1;
			start_elems = trace_ptr[inet + 1];
		}
	}
	*bounds_out = b;
	return nc;
}

typedef struct {
	const pf_problem *p; const pf_names *n; const pf_result *r;
	const int *bounds;
	outbuf *bufs;
	par_err err;
} write_ctx;

static void write_chunk(void *vctx, int k) {
	write_ctx *c = (write_ctx *)vctx;
	int lo = c->bounds[k], hi = c->bounds[k + 1], rc;
	size_t est = (size_t)(c->r->trace_ptr[hi] - c->r->trace_ptr[lo]) * 56 + (size_t)(hi - lo) * 48 + 4096;
	if (c->err.rc) { c->bufs[k].buf = NULL; c->bufs[k].len = 0; return; }
	if (ob_mem_open(&c->bufs[k], est) != 0) { c->bufs[k].len = 0; fail(PF_ENOMEM, "out of memory"); par_fail(&c->err, PF_ENOMEM); return; }
	rc = format_nets(&c->bufs[k], c->p, c->n, c->r, lo, hi);
	if (rc) par_fail(&c->err, rc);
}

int pf_route_write(const char *path, const pf_problem *p, const pf_names *n, const pf_result *r) {
	outbuf o;
	int rc, k, nchunks, *bounds = NULL;
	write_ctx c;
	if (!path || !p || !n || !r) return fail(PF_EINVAL, "pf_route_write: NULL argument");
	if (n->num_nets != p->num_nets || r->num_nets != p->num_nets || n->nx != p->nx || n->ny != p->ny)
		return fail(PF_EINVAL, "pf_route_write: problem, names and result disagree on nets or grid");
	if (!r->trace_ptr || (r->trace_ptr[p->num_nets] > 0 && !r->trace_node) || !n->net_name_ptr || !n->tile_is_io || !n->gpin_ptr
			|| !p->net_ptr || !p->net_is_global || !p->type || !p->xlow || !p->ylow || !p->xhigh || !p->yhigh || !p->ptc_num)
		return fail(PF_EINVAL, "pf_route_write: a required array is NULL");
	for (k = 0; k < p->num_nets; k++)
		if (r->trace_ptr[k + 1] < r->trace_ptr[k]) return fail(PF_EINVAL, "pf_route_write: trace_ptr not monotone at net %d", k);
	if ((rc = ob_open(&o, path)) != 0) return fail(rc, "cannot open %s for writing", path);
	ob_lit(&o, "Array size: "); ob_int(&o, p->nx); ob_lit(&o, " x "); ob_int(&o, p->ny); ob_lit(&o, " logic blocks.\n");
	ob_lit(&o, "\nRouting:");
	nchunks = p->num_nets > 0 ? chunk_nets(r->trace_ptr, p->num_nets, 1L << 16, &bounds) : 0;
	if (nchunks < 0) { rc = PF_ENOMEM; goto out; }
	if (nchunks <= 1 || par_threads() == 1) {                   /* small routing: straight into the file window */
		rc = format_nets(&o, p, n, r, 0, p->num_nets);
		goto out;
	}
	memset(&c, 0, sizeof(c));
	c.p = p; c.n = n; c.r = r; c.bounds = bounds;
	c.bufs = (outbuf *)calloc((size_t)nchunks, sizeof(outbuf));
	if (!c.bufs) { rc = PF_ENOMEM; goto out; }
	par_for(nchunks, write_chunk, &c);
	if (c.err.rc) { rc = c.err.rc; memcpy(g_err, c.err.msg, sizeof(g_err)); }
	ob_flush(&o);                                                /* the header precedes the chunks */
	for (k = 0; k < nchunks; k++) {
		if (!rc && c.bufs[k].len && fwrite(c.bufs[k].buf, 1, c.bufs[k].len, o.f) != c.bufs[k].len) o.err = 1;
		free(c.bufs[k].buf);
	}
	free(c.bufs);
out:
	free(bounds);
	k = ob_close(&o);
	return rc ? rc : k;
}

/* whole file into memory, NUL-terminated */
static int slurp(const char *path, char **data, size_t *len) {
	FILE *f = fopen(path, "rb");
	long sz;
	*data = NULL; *len = 0;
	if (!f) return fail(PF_EIO, "cannot open %s", path);
	if (fseek(f, 0, SEEK_END) != 0 || (sz = ftell(f)) < 0 || fseek(f, 0, SEEK_SET) != 0) { fclose(f); return fail(PF_EIO, "cannot size %s", path); }
	*data = (char *)malloc((size_t)sz + 1);
	if (!*data) { fclose(f); return PF_ENOMEM; }
	if (fread(*data, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(*data); *data = NULL; return fail(PF_EIO, "short read of %s", path); }
	fclose(f);
	(*data)[sz] = 0;
	*len = (size_t)sz;
	return PF_OK;
}

/* decimal integer at *s (optional '-'), advances *s; 0 if none */
static inline int scan_int(const char **s, long *v) {
	const char *c = *s;
	int neg = 0;
	long x = 0;
	if (*c == '-') { neg = 1; c++; }
	int digits = 0;
	if (*c < '0' || *c > '9') return 0;
	while (*c >= '0' && *c <= '9') { if (++digits > 18) return 0; x = x * 10 + (*c - '0'); c++; }   /* no signed overflow */
	*v = neg ? -x : x;
	*s = c;
	return 1;
}
static inline int expect(const char **s, const char *lit) {
	size_t n = strlen(lit);
	if (strncmp(*s, lit, n) != 0) return 0;
	*s += n;
	return 1;
}

typedef struct { int32_t *v; size_t n, cap; } ivec;
static int iv_push(ivec *a, int32_t x) {
	if (a->n == a->cap) {
		size_t nc = a->cap ? a->cap * 2 : 1 << 14;
		int32_t *nv = (int32_t *)realloc(a->v, nc * sizeof(int32_t));
		if (!nv) return PF_ENOMEM;
		a->v = nv; a->cap = nc;
	}
	a->v[a->n++] = x;
	return 0;
}

/* one piece of a .route file: begins at the start of the file or at a "Net <i> (" line */
typedef struct {
	const char *begin, *end;
	int first;               /* piece 0 carries the "Array size:" header */
	ivec nodes;              /* rr nodes of the piece's Node lines */
	ivec net_start;          /* for each net of the piece, in order: offset into nodes */
	int first_net;           /* index of the piece's first net, -1 if it has none */
	int rc, err_line;        /* err_line is local to the piece */
	char msg[200];
} route_piece;

typedef struct { const char *path; const pf_problem *p; route_piece *pieces; } read_ctx;

#define PIECE_FAIL(...) do { pc->rc = PF_EFORMAT; pc->err_line = line; snprintf(pc->msg, sizeof(pc->msg), __VA_ARGS__); return; } while (0)

static void parse_piece(void *vctx, int k) {
	read_ctx *ctx = (read_ctx *)vctx;
	route_piece *pc = &ctx->pieces[k];
	const pf_problem *p = ctx->p;
	const char *s = pc->begin, *end = pc->end;
	long a, b;
	int line = 0, cur = -1;
	pc->first_net = -1;
	while (s < end) {
		const char *eol = (const char *)memchr(s, '\n', (size_t)(end - s));
		const char *c = s;
		if (!eol) eol = end;
		line++;
		if (eol == s) { s = eol + 1; continue; }
		if (pc->first && line == 1) {
			if (!expect(&c, "Array size: ") || !scan_int(&c, &a) || !expect(&c, " x ") || !scan_int(&c, &b)) PIECE_FAIL("not a .route file (no 'Array size:' line)");
			if (a != p->nx || b != p->ny) PIECE_FAIL("routing of a %ld x %ld array, problem is %d x %d", a, b, p->nx, p->ny);
		} else if (c[0] == 'N' && c[1] == 'o') {                  /* "Node:\t<id>\t<type> (x,y) [to (x,y) ] <what>: <ptc>  " */
			long id, x, y, xh, yh, ptc;
			int t;
			if (cur < 0 || p->net_is_global[cur]) PIECE_FAIL("Node line outside a routed net");
			if (!expect(&c, "Node:\t") || !scan_int(&c, &id) || *c != '\t') PIECE_FAIL("malformed Node line");
			c++;
			while (*c == ' ') c++;
			for (t = 0; t < 6; t++) { size_t tl = strlen(TYPE_NAME[t]); if (strncmp(c, TYPE_NAME[t], tl) == 0 && c[tl] == ' ') { c += tl; break; } }
			if (t == 6 || !expect(&c, " (") || !scan_int(&c, &x) || !expect(&c, ",") || !scan_int(&c, &y) || !expect(&c, ") ")) PIECE_FAIL("malformed Node line");
			xh = x; yh = y;
			if (c[0] == 't' && (!expect(&c, "to (") || !scan_int(&c, &xh) || !expect(&c, ",") || !scan_int(&c, &yh) || !expect(&c, ") "))) PIECE_FAIL("malformed 'to (x,y)'");
			while (c < eol && *c != ':') c++;                        /* " Pad" / " Pin" / " Track" / " Class" */
			if (c >= eol || c[1] != ' ') PIECE_FAIL("malformed Node line");
			c += 2;
			if (!scan_int(&c, &ptc)) PIECE_FAIL("no track / pin / class number");
			if (id < 0 || id >= p->num_nodes) PIECE_FAIL("rr node %ld out of range", id);
			if (p->type[id] != t || p->xlow[id] != x || p->ylow[id] != y || p->xhigh[id] != xh || p->yhigh[id] != yh || p->ptc_num[id] != ptc)
				PIECE_FAIL("rr node %ld is %s (%d,%d)-(%d,%d) ptc %d in the problem: the file belongs to another rr graph",
						id, TYPE_NAME[p->type[id] <= PF_CHANY ? p->type[id] : 0], p->xlow[id], p->ylow[id], p->xhigh[id], p->yhigh[id], p->ptc_num[id]);
			if (iv_push(&pc->nodes, (int32_t)id) != 0) { pc->rc = PF_ENOMEM; return; }
		} else if (c[0] == 'N' && c[1] == 'e') {                  /* "Net <i> (<name>)" or "...): global net connecting:" */
			static const char GLOB[] = "): global net connecting:";      /* matched at the tail: names may contain ')' */
			const size_t gl = sizeof(GLOB) - 1;
			int glob;
			if (!expect(&c, "Net ") || !scan_int(&c, &a) || !expect(&c, " (")) PIECE_FAIL("malformed Net line");
			if (a < 0 || a >= p->num_nets || (cur >= 0 && a != cur + 1) || (pc->first && cur < 0 && a != 0))
				PIECE_FAIL("net %ld follows net %d (problem has %d nets)", a, cur, p->num_nets);
			cur = (int)a;
			if (pc->first_net < 0) pc->first_net = cur;
			if (iv_push(&pc->net_start, (int32_t)pc->nodes.n) != 0) { pc->rc = PF_ENOMEM; return; }
			glob = (size_t)(eol - c) >= gl && memcmp(eol - gl, GLOB, gl) == 0;
			if (!glob && eol[-1] != ')') PIECE_FAIL("malformed Net line");
			if (glob != (p->net_is_global[cur] != 0)) PIECE_FAIL("net %d global in one of file / problem only", cur);
		} else if (strncmp(c, "Routing:", 8) == 0 || strncmp(c, "Block ", 6) == 0 || strncmp(c, "Used in local", 13) == 0) {
			/* nothing to take from these */
		} else {
			PIECE_FAIL("unrecognised line");
		}
		s = eol + 1;
	}
}

/* per net: the switch of every trace element, the element's serial-number term, the net's wirelength */
typedef struct {
	const char *path; const pf_problem *p;
	const int *bounds;
	const int32_t *tptr, *nodes;
	int16_t *tsw;
	uint32_t *term;
	long *chunk_wl;
	par_err err;
} finish_ctx;

static void finish_chunk(void *vctx, int k) {
	finish_ctx *c = (finish_ctx *)vctx;
	const pf_problem *p = c->p;
	int inet;
	long wl = 0;
	for (inet = c->bounds[k]; inet < c->bounds[k + 1]; inet++) {
		size_t lo = (size_t)c->tptr[inet], hi = (size_t)c->tptr[inet + 1], i;
		const uint32_t mult = (uint32_t)(inet + 1);
		for (i = lo; i < hi; i++) {
			int u = c->nodes[i], e, found = -1;
			/* get_serial_num, route_common.c:224-254, in the reference's wrapping int arithmetic */
			c->term[i] = mult * (uint32_t)(p->xlow[u] * (p->nx + 1) - p->yhigh[u] - 10 * (int)p->ptc_num[u] - 100 * (int)p->type[u]);
			if (p->type[u] == PF_SINK) { c->tsw[i] = PF_OPEN; continue; }
			if (i + 1 >= hi) { fail(PF_EFORMAT, "%s: net %d does not end at a SINK", c->path, inet); par_fail(&c->err, PF_EFORMAT); return; }
			for (e = p->row_ptr[u]; e < p->row_ptr[u + 1]; e++)
				if (p->edge_to[e] == c->nodes[i + 1]) { found = e; break; }
			if (found < 0) { fail(PF_EFORMAT, "%s: net %d: no rr edge %d -> %d", c->path, inet, u, c->nodes[i + 1]); par_fail(&c->err, PF_EFORMAT); return; }
			c->tsw[i] = p->edge_sw[found];
		}
		/* get_num_bends_and_length, base/stats.c:355-409: the first element and every element after a SINK are skipped */
		for (i = lo + 1; i < hi; i++) {
			int u = c->nodes[i], t = p->type[u];
			if (t == PF_SINK) { i++; continue; }
			if (t == PF_CHANX || t == PF_CHANY) wl += 1 + p->xhigh[u] - p->xlow[u] + p->yhigh[u] - p->ylow[u];
		}
	}
	c->chunk_wl[k] = wl;
}

int pf_route_read(const char *path, const pf_problem *p, pf_result *r) {
	char *data = NULL;
	size_t len = 0, total = 0, i;
	int rc, k, npieces, nets_seen = 0, nchunks = 0, *bounds = NULL;
	route_piece *pieces = NULL;
	int32_t *tptr = NULL, *nodes = NULL;
	int16_t *tsw = NULL;
	uint32_t *term = NULL;
	long *chunk_wl = NULL;
	read_ctx rctx;
	finish_ctx fctx;
	memset(r, 0, sizeof(*r));
	g_err[0] = 0;
	if ((rc = slurp(path, &data, &len)) != 0) return rc;

	/* pieces of >= 4 MB, cut where a "Net " header follows a blank line */
	npieces = par_threads() * 4;
	if ((size_t)npieces > len / ((size_t)4 << 20) + 1) npieces = (int)(len / ((size_t)4 << 20)) + 1;
	pieces = (route_piece *)calloc((size_t)npieces, sizeof(route_piece));
	if (!pieces) { rc = PF_ENOMEM; goto done; }
	{
		const char *at = data;
		int made = 0;
		for (k = 0; k < npieces; k++) {
			const char *cut = data + len;
			if (k + 1 < npieces) {
				const char *from = data + len / (size_t)npieces * (size_t)(k + 1);
				const char *hit = from > at ? strstr(from, "\n\nNet ") : NULL;
				if (hit) cut = hit + 2;
			}
			if (cut <= at) continue;
			pieces[made].begin = at; pieces[made].end = cut; pieces[made].first = made == 0;
			made++;
			at = cut;
			if (cut == data + len) break;
		}
		npieces = made;
	}
	rctx.path = path; rctx.p = p; rctx.pieces = pieces;
	if (npieces > 0) par_for(npieces, parse_piece, &rctx);
	for (k = 0; k < npieces; k++) {
		route_piece *pc = &pieces[k];
		if (pc->rc == PF_EFORMAT) {
			long line = pc->err_line;
			const char *c;
			for (c = data; c < pc->begin; c++) if (*c == '\n') line++;
			rc = fail(PF_EFORMAT, "%s:%ld: %s", path, line, pc->msg); goto done;
		}
		if (pc->rc) { rc = pc->rc; goto done; }
		if (pc->net_start.n && pc->first_net != nets_seen) { rc = fail(PF_EFORMAT, "%s: net %d follows net %d (problem has %d nets)", path, pc->first_net, nets_seen - 1, p->num_nets); goto done; }
		nets_seen += (int)pc->net_start.n;
		total += pc->nodes.n;
	}
	if (nets_seen != p->num_nets) { rc = fail(PF_EFORMAT, "%s: %d nets in the file, %d in the problem", path, nets_seen, p->num_nets); goto done; }
	if (total > 0x7fffffffu) { rc = fail(PF_EFORMAT, "%s: more than 2^31 trace elements", path); goto done; }
	tptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)p->num_nets + 1));
	nodes = (int32_t *)malloc(sizeof(int32_t) * (total ? total : 1));
	tsw = (int16_t *)malloc(sizeof(int16_t) * (total ? total : 1));
	term = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));
	if (!tptr || !nodes || !tsw || !term) { rc = PF_ENOMEM; goto done; }
	{
		size_t base = 0;
		int inet = 0;
		for (k = 0; k < npieces; k++) {
			route_piece *pc = &pieces[k];
			for (i = 0; i < pc->net_start.n; i++) tptr[inet++] = (int32_t)(base + (size_t)pc->net_start.v[i]);
			if (pc->nodes.n) memcpy(nodes + base, pc->nodes.v, sizeof(int32_t) * pc->nodes.n);
			base += pc->nodes.n;
			free(pc->nodes.v); pc->nodes.v = NULL;
			free(pc->net_start.v); pc->net_start.v = NULL;
		}
		tptr[p->num_nets] = (int32_t)total;
	}
	nchunks = p->num_nets > 0 ? chunk_nets(tptr, p->num_nets, 1L << 16, &bounds) : 0;
	if (nchunks < 0) { rc = PF_ENOMEM; goto done; }
	chunk_wl = (long *)calloc((size_t)(nchunks ? nchunks : 1), sizeof(long));
	if (!chunk_wl) { rc = PF_ENOMEM; goto done; }
	memset(&fctx, 0, sizeof(fctx));
	fctx.path = path; fctx.p = p; fctx.bounds = bounds; fctx.tptr = tptr; fctx.nodes = nodes; fctx.tsw = tsw; fctx.term = term; fctx.chunk_wl = chunk_wl;
	if (nchunks > 0) par_for(nchunks, finish_chunk, &fctx);
	if (fctx.err.rc) { rc = fctx.err.rc; memcpy(g_err, fctx.err.msg, sizeof(g_err)); goto done; }
	{
		/* the running remainder is sequential by definition: serial = (serial + term) % 2000000000 with C's truncating
		 * remainder; |x| < 2^31 < 2 * 2000000000, so it is one conditional add or subtract */
		const int M = 2000000000;
		int sv = 0;
		long wl = 0;
		for (i = 0; i < total; i++) {
			int x = (int)((uint32_t)sv + term[i]);
			sv = x >= M ? x - M : (x <= -M ? x + M : x);
		}
		for (k = 0; k < nchunks; k++) wl += chunk_wl[k];
		r->serial_num = sv;
		r->total_wirelength = (int32_t)wl;
	}
	r->num_nets = p->num_nets;
	r->trace_ptr = tptr; tptr = NULL;
	r->trace_node = nodes; nodes = NULL;
	r->trace_switch = tsw; tsw = NULL;
	rc = PF_OK;
done:
	if (pieces) for (k = 0; k < npieces; k++) { free(pieces[k].nodes.v); free(pieces[k].net_start.v); }
	free(pieces); free(data); free(tptr); free(nodes); free(tsw); free(term); free(chunk_wl); free(bounds);
	return rc;
}

/* ------------------------------------------------------------------ .place */
int pf_place_write(const char *path, const char *net_file, const char *arch_file, const pf_names *n) {
	outbuf o;
	int i, rc;
	if (!path || !net_file || !arch_file || !n) return fail(PF_EINVAL, "pf_place_write: NULL argument");
	if ((rc = ob_open(&o, path)) != 0) return rc;
	ob_lit(&o, "Netlist file: "); ob_mem(&o, net_file, strlen(net_file));
	ob_lit(&o, "   Architecture file: "); ob_mem(&o, arch_file, strlen(arch_file)); ob_lit(&o, "\n");
	ob_lit(&o, "Array size: "); ob_int(&o, n->nx); ob_lit(&o, " x "); ob_int(&o, n->ny); ob_lit(&o, " logic blocks\n\n");
	ob_lit(&o, "#block name\tx\ty\tsubblk\tblock number\n");
	ob_lit(&o, "#----------\t--\t--\t------\t------------\n");
	for (i = 0; i < n->num_blocks; i++) {
		size_t nl = (size_t)(n->block_name_ptr[i + 1] - n->block_name_ptr[i]);
		ob_mem(&o, n->block_name_chars + n->block_name_ptr[i], nl); ob_lit(&o, "\t");
		if (nl < 8) ob_lit(&o, "\t");
		ob_int(&o, n->block_x[i]); ob_lit(&o, "\t"); ob_int(&o, n->block_y[i]); ob_lit(&o, "\t"); ob_int(&o, n->block_z[i]);
		ob_lit(&o, "\t#"); ob_int(&o, i); ob_lit(&o, "\n");
	}
	return ob_close(&o);
}

static uint64_t fnv1a(const char *s, size_t n) {
	uint64_t h = 1469598103934665603ull;
	size_t i;
	for (i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
	return h;
}

/* One logical line as ReadLineTokens (libarchfpga/ReadLine.c:38-196) sees it: '\r' stripped, a trailing backslash
 * joins the next physical line, '#' starts a comment, tokens split at blanks and tabs; lines without tokens are
 * skipped.  Tokenises the buffer in place; returns the token count (capped at MAX_TOK), -1 at end of file. */
#define MAX_TOK 16
static int next_tokens(char **cursor, char *end, int *line, char *tok[MAX_TOK], size_t tok_len[MAX_TOK]) {
	for (;;) {
		char *s = *cursor, *start, *w, *c, *last;
		int nt = 0;
		if (s >= end) return -1;
		start = w = s;                           /* continued lines are compacted towards the front */
		for (;;) {
			char *eol = (char *)memchr(s, '\n', (size_t)(end - s));
			char *stop = eol ? eol : end;
			char *next = eol ? eol + 1 : end;
			(*line)++;
			if (stop > s && stop[-1] == '\r') stop--;
			if (stop > s && stop[-1] == '\\' && next < end) {
				memmove(w, s, (size_t)(stop - 1 - s)); w += stop - 1 - s;
				s = next;
				continue;
			}
			memmove(w, s, (size_t)(stop - s)); w += stop - s;
			*cursor = next;
			break;
		}
		last = w;
		c = (char *)memchr(start, '#', (size_t)(last - start));
		if (c) last = c;
		c = start;
		while (c < last) {
			char *b;
			while (c < last && (*c == ' ' || *c == '\t')) c++;
			if (c >= last) break;
			b = c;
			while (c < last && *c != ' ' && *c != '\t') c++;
			if (nt < MAX_TOK) { tok[nt] = b; tok_len[nt] = (size_t)(c - b); }
			nt++;
		}
		if (nt > 0) return nt > MAX_TOK ? MAX_TOK : nt;
	}
}

static int tok_is(const char *t, size_t n, const char *lit) { return strlen(lit) == n && memcmp(t, lit, n) == 0; }
/* my_atoi (util.c): the token must start with a digit or '-' */
static int tok_int(const char *t, size_t n, long *v) {
	const char *c = t;
	if (n == 0 || !scan_int(&c, v)) return 0;
	return 1;
}

int pf_place_read(const char *path, const char *net_file, const char *arch_file, pf_names *n, int *placed) {
	char *data = NULL, *cur, *end;
	size_t len = 0, cap, mask, h;
	int32_t *table = NULL;
	char *tok[MAX_TOK];
	size_t tl[MAX_TOK];
	int line = 0, nt, i, rc, count = 0;
	long a, b, z;
	g_err[0] = 0;
	if (placed) *placed = 0;
	if ((rc = slurp(path, &data, &len)) != 0) return rc;
	cur = data; end = data + len;
	/* "Netlist file: <net>   Architecture file: <arch>"  (read_place.c:29-66) */
	nt = next_tokens(&cur, end, &line, tok, tl);
	if (nt < 6 || !tok_is(tok[0], tl[0], "Netlist") || !tok_is(tok[1], tl[1], "file:") || !tok_is(tok[3], tl[3], "Architecture")
			|| !tok_is(tok[4], tl[4], "file:")) {
		rc = fail(PF_EFORMAT, "'%s' - Bad filename specification line in placement file.", path); goto done;
	}
	if (arch_file && !tok_is(tok[5], tl[5], arch_file)) {
		rc = fail(PF_EFORMAT, "'%s' - Architecture file that generated placement (%.*s) does not match current architecture file (%s).",
				path, (int)tl[5], tok[5], arch_file);
		goto done;
	}
	if (net_file && !tok_is(tok[2], tl[2], net_file)) {
		rc = fail(PF_EFORMAT, "'%s' - Netlist file that generated placement (%.*s) does not match current netlist file (%s).",
				path, (int)tl[2], tok[2], net_file);
		goto done;
	}
	/* "Array size: <nx> x <ny> logic blocks"  (read_place.c:68-103) */
	nt = next_tokens(&cur, end, &line, tok, tl);
	if (nt < 7 || !tok_is(tok[0], tl[0], "Array") || !tok_is(tok[1], tl[1], "size:") || !tok_is(tok[3], tl[3], "x")
			|| !tok_is(tok[5], tl[5], "logic") || !tok_is(tok[6], tl[6], "blocks") || !tok_int(tok[2], tl[2], &a) || !tok_int(tok[4], tl[4], &b)) {
		rc = fail(PF_EFORMAT, "'%s' - Bad FPGA size specification line in placement file.", path); goto done;
	}
	if (a != n->nx || b != n->ny) {
		rc = fail(PF_EFORMAT, "'%s' - Current FPGA size (%d x %d) is different from size when placement generated (%ld x %ld).", path, n->nx, n->ny, a, b);
		goto done;
	}
	/* block name -> index; the first block of a name wins, as in the reference's linear scan */
	cap = 16;
	while (cap < (size_t)n->num_blocks * 2) cap <<= 1;
	mask = cap - 1;
	table = (int32_t *)malloc(cap * sizeof(int32_t));
	if (!table) { rc = PF_ENOMEM; goto done; }
	memset(table, 0xff, cap * sizeof(int32_t));
	for (i = 0; i < n->num_blocks; i++) {
		const char *nm = n->block_name_chars + n->block_name_ptr[i];
		size_t nl = (size_t)(n->block_name_ptr[i + 1] - n->block_name_ptr[i]);
		int dup = 0;
		for (h = fnv1a(nm, nl) & mask; table[h] >= 0; h = (h + 1) & mask) {
			int j = table[h];
			if ((size_t)(n->block_name_ptr[j + 1] - n->block_name_ptr[j]) == nl && memcmp(n->block_name_chars + n->block_name_ptr[j], nm, nl) == 0) { dup = 1; break; }
		}
		if (!dup) table[h] = i;
	}
	while ((nt = next_tokens(&cur, end, &line, tok, tl)) >= 0) {           /* read_place.c:105-135 */
		int blk = -1;
		for (h = fnv1a(tok[0], tl[0]) & mask; table[h] >= 0; h = (h + 1) & mask) {
			int j = table[h];
			if ((size_t)(n->block_name_ptr[j + 1] - n->block_name_ptr[j]) == tl[0] && memcmp(n->block_name_chars + n->block_name_ptr[j], tok[0], tl[0]) == 0) { blk = j; break; }
		}
		if (blk < 0) { rc = fail(PF_EFORMAT, "'%s':%d - Block in placement file does not exist in netlist.", path, line); goto done; }
		if (nt < 4 || !tok_int(tok[1], tl[1], &a) || !tok_int(tok[2], tl[2], &b) || !tok_int(tok[3], tl[3], &z)) {
			rc = fail(PF_EFORMAT, "'%s':%d - expected <block> <x> <y> <subblk>.", path, line); goto done;
		}
		n->block_x[blk] = (int32_t)a; n->block_y[blk] = (int32_t)b; n->block_z[blk] = (int32_t)z;
		count++;
	}
	if (placed) *placed = count;
	rc = PF_OK;
done:
	free(data); free(table);
	return rc;
}

/* ------------------------------------------------------------------ pf_net_read: the packed netlist (.net)
 * What read_netlist (vpr/SRC/base/read_netlist.c:74-244) leaves behind for place and route: block[] (name, type, the net on
 * every pin) and clb_net[] (name, driver + sinks as (block, pin), is_global).  The reference instantiates the architecture's
 * pb_type hierarchy to get there (processComplexBlock :246, processPb :362, processPorts :610) and reads the nets on a
 * cluster's output pins out of the cluster's internal rr graph (load_external_nets_and_cb :836-984).  Here the file is its
 * own description: the <inputs> / <outputs> / <clocks> sections of a top-level block list every port in pb_type order with
 * "open" for unused pins (output_clustering.c), which IS the pin numbering (inputs, outputs, clocks — the order
 * load_external_nets_and_cb asserts, :850), and an output pin "child[i].port[b]->interconnect" is followed down the nested
 * <block> elements to the primitive whose port holds the net's name. */
typedef struct { const char *s; int n; } nl_str;
typedef struct { int pb, kind; nl_str name; int first_tok, ntok; } nl_port;
typedef struct { int parent, first_child, last_child, next_sibling; nl_str name, type; int index, first_port, nports, line; } nl_pb;
typedef struct {
	nl_pb *pb; int npb, cap_pb;
	nl_port *port; int nport, cap_port;
	nl_str *tok; int ntok, cap_tok;
	nl_str *clk; int nclk, cap_clk;          /* <clocks> of the FPGA_packed_netlist element */
} nl_doc;

static int nl_grow(void **v, int *cap, int need, size_t elem) {
	if (need <= *cap) return 0;
	int nc = *cap ? *cap * 2 : 1024;
	while (nc < need) nc *= 2;
	void *nv = realloc(*v, (size_t)nc * elem);
	if (!nv) return PF_ENOMEM;
	*v = nv; *cap = nc;
	return 0;
}
static int nl_eq(nl_str a, const char *s, int n) { return a.n == n && memcmp(a.s, s, (size_t)n) == 0; }
static int nl_is_open(nl_str a) { return nl_eq(a, "open", 4); }
/* attribute value inside a tag: name="value" */
static int nl_attr(const char *tag, const char *tag_end, const char *name, nl_str *out) {
	size_t ln = strlen(name);
	const char *c = tag;
	while (c + ln + 2 < tag_end) {
		if ((c == tag || c[-1] == ' ' || c[-1] == '\t' || c[-1] == '\n') && memcmp(c, name, ln) == 0 && c[ln] == '=' && c[ln + 1] == '"') {
			const char *v = c + ln + 2, *e = (const char *)memchr(v, '"', (size_t)(tag_end - v));
			if (!e) return 0;
			out->s = v; out->n = (int)(e - v);
			return 1;
		}
		c++;
	}
	return 0;
}
/* "type[index]" */
static int nl_instance(nl_str in, nl_str *type, int *index) {
	int k = in.n - 1;
	long v;
	const char *c;
	if (k < 2 || in.s[k] != ']') return 0;
	while (k > 0 && in.s[k] != '[') k--;
	if (k <= 0) return 0;
	c = in.s + k + 1;
	if (!scan_int(&c, &v) || c != in.s + in.n - 1 || v < 0 || v > 0x7fffffff) return 0;
	type->s = in.s; type->n = k; *index = (int)v;
	return 1;
}
static nl_port *nl_find_port(nl_doc *d, int pb, int want_out, const char *name, int n) {
	for (int k = 0; k < d->pb[pb].nports; k++) {
		nl_port *q = &d->port[d->pb[pb].first_port + k];
		if ((q->kind == 1) == (want_out != 0) && nl_eq(q->name, name, n)) return q;
	}
	return NULL;
}
/* The net on the pin that `t` names, read in the context of block `ctx` (the element whose port, or whose child's port, holds
 * the token): "child[i].port[b]->ic" is an output of a child of ctx; "type.port[b]->ic" is an input of ctx itself, whose own
 * token is read in the context of ctx's parent.  A token without "->" is a net name (primitive outputs, cluster inputs). */
static int nl_resolve(nl_doc *d, int ctx, nl_str t, nl_str *net, int depth, int *err_line) {
	for (;;) {
		const char *arrow = NULL, *dot, *br, *c;
		long bit;
		nl_str inst, type, pname;
		int index = -1, has_index;
		nl_port *q;
		for (int k = 0; k + 1 < t.n; k++) if (t.s[k] == '-' && t.s[k + 1] == '>') { arrow = t.s + k; break; }
		if (!arrow) { *net = t; return 0; }
		if (++depth > 256) { *err_line = d->pb[ctx].line; return -1; }
		dot = (const char *)memchr(t.s, '.', (size_t)(arrow - t.s));
		if (!dot) { *err_line = d->pb[ctx].line; return -1; }
		inst.s = t.s; inst.n = (int)(dot - t.s);
		br = (const char *)memchr(dot, '[', (size_t)(arrow - dot));
		if (!br) { *err_line = d->pb[ctx].line; return -1; }
		pname.s = dot + 1; pname.n = (int)(br - dot - 1);
		c = br + 1;
		if (!scan_int(&c, &bit) || *c != ']' || bit < 0) { *err_line = d->pb[ctx].line; return -1; }
		has_index = nl_instance(inst, &type, &index);
		if (has_index) {
			int ch = d->pb[ctx].first_child;
			while (ch >= 0 && !(d->pb[ch].index == index && nl_eq(d->pb[ch].type, type.s, type.n))) ch = d->pb[ch].next_sibling;
			if (ch < 0) { *err_line = d->pb[ctx].line; return -1; }
			q = nl_find_port(d, ch, 1, pname.s, pname.n);
			if (!q || bit >= q->ntok) { *err_line = d->pb[ch].line; return -1; }
			t = d->tok[q->first_tok + (int)bit];
			ctx = ch;
		} else {
			if (!nl_eq(d->pb[ctx].type, inst.s, inst.n)) { *err_line = d->pb[ctx].line; return -1; }
			q = nl_find_port(d, ctx, 0, pname.s, pname.n);
			if (!q || bit >= q->ntok) { *err_line = d->pb[ctx].line; return -1; }
			t = d->tok[q->first_tok + (int)bit];
			if (d->pb[ctx].parent <= 0) { *net = t; return 0; }      /* a cluster input: the token is the net */
			ctx = d->pb[ctx].parent;
		}
	}
}

void pf_netlist_free(pf_netlist *nl) {
	if (!nl) return;
	free(nl->block_name_ptr); free(nl->block_name_chars); free(nl->block_type_ptr); free(nl->block_type_chars);
	free(nl->block_pin_ptr); free(nl->block_pin_net); free(nl->block_pin_kind);
	free(nl->net_name_ptr); free(nl->net_name_chars); free(nl->net_ptr); free(nl->net_block); free(nl->net_block_pin); free(nl->net_is_global);
	memset(nl, 0, sizeof(*nl));
}

int pf_net_read(const char *path, pf_netlist *out) {
	char *data = NULL;
	size_t len = 0;
	nl_doc d;
	int rc, line = 1, sp = 0, cap_stack = 0, *stack = NULL, cur_kind = -1, cur_port = -1, root = -1;
	const char *p, *end;
	int32_t *pin_net = NULL, *table = NULL, *count = NULL, *fill = NULL;
	nl_str *net_names = NULL;
	int nnets = 0, cap_nets = 0, nblocks = 0, npins = 0, err_line = 0;
	size_t tcap = 0, tmask = 0;
	g_err[0] = 0;
	memset(&d, 0, sizeof(d));
	if (!path || !out) return fail(PF_EINVAL, "null argument");
	memset(out, 0, sizeof(*out));
	if ((rc = slurp(path, &data, &len)) != PF_OK) return rc;
	p = data; end = data + len;
#define NL_FAIL(code, ...) do { rc = fail(code, __VA_ARGS__); goto done; } while (0)
	/* ---- pass 1: the element tree */
	while (p < end) {
		if (*p == '\n') { line++; p++; continue; }
		if (*p == ' ' || *p == '\t' || *p == '\r') { p++; continue; }
		if (*p == '<') {
			const char *te = (const char *)memchr(p, '>', (size_t)(end - p)), *nm = p + 1, *ne;
			int closing = 0, selfclose;
			if (!te) NL_FAIL(PF_EFORMAT, "%s:%d: unterminated tag", path, line);
			if (p[1] == '?' || p[1] == '!') { for (const char *c = p; c < te; c++) if (*c == '\n') line++; p = te + 1; continue; }
			if (*nm == '/') { closing = 1; nm++; }
			selfclose = te > p && te[-1] == '/';
			ne = nm;
			while (ne < te && *ne != ' ' && *ne != '\t' && *ne != '\n' && *ne != '/' ) ne++;
			int tl = (int)(ne - nm);
			if (tl == 5 && memcmp(nm, "block", 5) == 0) {
				if (closing) {
					if (sp == 0) NL_FAIL(PF_EFORMAT, "%s:%d: </block> without <block>", path, line);
					sp--;
				} else {
					nl_str name, inst;
					nl_pb *b;
					if (!nl_attr(ne, te, "name", &name) || !nl_attr(ne, te, "instance", &inst)) NL_FAIL(PF_EFORMAT, "%s:%d: <block> needs name and instance", path, line);
					if (nl_grow((void **)&d.pb, &d.cap_pb, d.npb + 1, sizeof(nl_pb))) NL_FAIL(PF_ENOMEM, "out of memory");
					b = &d.pb[d.npb];
					memset(b, 0, sizeof(*b));
					b->parent = sp ? stack[sp - 1] : -1; b->first_child = b->last_child = b->next_sibling = -1;
					b->name = name; b->line = line; b->first_port = d.nport;
					if (!nl_instance(inst, &b->type, &b->index)) NL_FAIL(PF_EFORMAT, "%s:%d: instance \"%.*s\" is not type[index]", path, line, inst.n, inst.s);
					if (sp == 0) {
						/* read_netlist.c:101-108 */
						if (root >= 0) NL_FAIL(PF_EFORMAT, "%s:%d: second top-level element", path, line);
						if (!nl_eq(inst, "FPGA_packed_netlist[0]", 22)) NL_FAIL(PF_EFORMAT, "[Line %d] Expected instance to be \"FPGA_packed_netlist[0]\", found %.*s.", line, inst.n, inst.s);
						root = d.npb;
					} else {
						nl_pb *par = &d.pb[stack[sp - 1]];
						if (par->last_child >= 0) d.pb[par->last_child].next_sibling = d.npb; else par->first_child = d.npb;
						par->last_child = d.npb;
					}
					if (!selfclose) {
						if (nl_grow((void **)&stack, &cap_stack, sp + 1, sizeof(int))) NL_FAIL(PF_ENOMEM, "out of memory");
						stack[sp++] = d.npb;
					}
					d.npb++;
				}
				cur_kind = -1; cur_port = -1;
			} else if ((tl == 6 && memcmp(nm, "inputs", 6) == 0) || (tl == 7 && memcmp(nm, "outputs", 7) == 0) || (tl == 6 && memcmp(nm, "clocks", 6) == 0)) {
				cur_kind = (closing || selfclose) ? -1 : (nm[0] == 'i' ? 0 : nm[0] == 'o' ? 1 : 2);
				cur_port = -1;
			} else if (tl == 4 && memcmp(nm, "port", 4) == 0) {
				if (closing) cur_port = -1;
				else {
					nl_port *q;
					if (sp == 0 || cur_kind < 0) NL_FAIL(PF_EFORMAT, "%s:%d: <port> outside <inputs> / <outputs> / <clocks>", path, line);
					if (nl_grow((void **)&d.port, &d.cap_port, d.nport + 1, sizeof(nl_port))) NL_FAIL(PF_ENOMEM, "out of memory");
					q = &d.port[d.nport];
					q->pb = stack[sp - 1]; q->kind = cur_kind; q->first_tok = d.ntok; q->ntok = 0;
					if (!nl_attr(ne, te, "name", &q->name)) NL_FAIL(PF_EFORMAT, "%s:%d: <port> needs a name", path, line);
					if (d.pb[q->pb].first_port + d.pb[q->pb].nports != d.nport) NL_FAIL(PF_EFORMAT, "%s:%d: ports of block \"%.*s\" are interleaved with a child block", path, line, d.pb[q->pb].name.n, d.pb[q->pb].name.s);
					d.pb[q->pb].nports++;
					cur_port = selfclose ? -1 : d.nport;
					d.nport++;
				}
			}
			for (const char *c = p; c < te; c++) if (*c == '\n') line++;
			p = te + 1;
			continue;
		}
		{	/* a text token */
			const char *b = p;
			nl_str t;
			while (p < end && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r' && *p != '<') p++;
			t.s = b; t.n = (int)(p - b);
			if (cur_port >= 0) {
				if (nl_grow((void **)&d.tok, &d.cap_tok, d.ntok + 1, sizeof(nl_str))) NL_FAIL(PF_ENOMEM, "out of memory");
				d.tok[d.ntok++] = t; d.port[cur_port].ntok++;
			} else if (sp == 1 && cur_kind == 2) {
				if (nl_grow((void **)&d.clk, &d.cap_clk, d.nclk + 1, sizeof(nl_str))) NL_FAIL(PF_ENOMEM, "out of memory");
				d.clk[d.nclk++] = t;
			}
		}
	}
	if (root < 0) NL_FAIL(PF_EFORMAT, "%s: no FPGA_packed_netlist element", path);
	if (sp != 0) NL_FAIL(PF_EFORMAT, "%s: <block> \"%.*s\" (line %d) is not closed", path, d.pb[stack[sp - 1]].name.n, d.pb[stack[sp - 1]].name.s, d.pb[stack[sp - 1]].line);
	/* ---- pass 2: the net on every pin of every complex block, in pin order (inputs, outputs, clocks) */
	for (int b = d.pb[root].first_child; b >= 0; b = d.pb[b].next_sibling) {
		nblocks++;
		for (int k = 0; k < d.pb[b].nports; k++) npins += d.port[d.pb[b].first_port + k].ntok;
	}
	out->num_blocks = nblocks;
	out->block_name_ptr = (int32_t *)calloc((size_t)nblocks + 1, sizeof(int32_t)); out->block_type_ptr = (int32_t *)calloc((size_t)nblocks + 1, sizeof(int32_t));
	out->block_pin_ptr = (int32_t *)calloc((size_t)nblocks + 1, sizeof(int32_t));
	out->block_pin_net = (int32_t *)malloc(sizeof(int32_t) * (size_t)(npins > 0 ? npins : 1)); out->block_pin_kind = (uint8_t *)malloc((size_t)(npins > 0 ? npins : 1));
	for (tcap = 1024; tcap < 4 * (size_t)npins + 16; tcap *= 2) {}
	tmask = tcap - 1;
	table = (int32_t *)malloc(sizeof(int32_t) * tcap);
	if (!out->block_name_ptr || !out->block_type_ptr || !out->block_pin_ptr || !out->block_pin_net || !out->block_pin_kind || !table) NL_FAIL(PF_ENOMEM, "out of memory");
	memset(table, 0xff, sizeof(int32_t) * tcap);
	pin_net = out->block_pin_net;
	{
		int ib = 0, at = 0;
		size_t name_bytes = 0, type_bytes = 0;
		for (int b = d.pb[root].first_child; b >= 0; b = d.pb[b].next_sibling, ib++) {
			name_bytes += (size_t)d.pb[b].name.n; type_bytes += (size_t)d.pb[b].type.n;
			out->block_name_ptr[ib + 1] = (int32_t)name_bytes; out->block_type_ptr[ib + 1] = (int32_t)type_bytes;
			for (int kind = 0; kind < 3; kind++)
				for (int k = 0; k < d.pb[b].nports; k++) {
					nl_port *q = &d.port[d.pb[b].first_port + k];
					if (q->kind != kind) continue;
					for (int j = 0; j < q->ntok; j++, at++) {
						nl_str net = d.tok[q->first_tok + j];
						out->block_pin_kind[at] = (uint8_t)kind;
						if (kind == 1 && nl_resolve(&d, b, net, &net, 0, &err_line) != 0)
							NL_FAIL(PF_EFORMAT, "%s:%d: cannot follow output %.*s[%d] of block \"%.*s\" to a primitive", path, err_line, q->name.n, q->name.s, j, d.pb[b].name.n, d.pb[b].name.s);
						if (nl_is_open(net)) { pin_net[at] = PF_OPEN; continue; }      /* add_net_to_hash :594: "open" is a keyword */
						size_t h = (size_t)fnv1a(net.s, (size_t)net.n) & tmask;
						while (table[h] >= 0 && !nl_eq(net_names[table[h]], net.s, net.n)) h = (h + 1) & tmask;
						if (table[h] < 0) {
							if (nl_grow((void **)&net_names, &cap_nets, nnets + 1, sizeof(nl_str))) NL_FAIL(PF_ENOMEM, "out of memory");
							net_names[nnets] = net; table[h] = nnets++;      /* index = order of first appearance (add_net_to_hash) */
						}
						pin_net[at] = table[h];
					}
				}
			out->block_pin_ptr[ib + 1] = at;
		}
		out->block_name_chars = (char *)malloc(name_bytes + 1); out->block_type_chars = (char *)malloc(type_bytes + 1);
		if (!out->block_name_chars || !out->block_type_chars) NL_FAIL(PF_ENOMEM, "out of memory");
		ib = 0;
		for (int b = d.pb[root].first_child; b >= 0; b = d.pb[b].next_sibling, ib++) {
			memcpy(out->block_name_chars + out->block_name_ptr[ib], d.pb[b].name.s, (size_t)d.pb[b].name.n);
			memcpy(out->block_type_chars + out->block_type_ptr[ib], d.pb[b].type.s, (size_t)d.pb[b].type.n);
		}
	}
	/* ---- pass 3: clb_net[] — the driver first, the sinks in block / pin order (load_external_nets_and_cb :934-965) */
	out->num_nets = nnets;
	out->net_name_ptr = (int32_t *)calloc((size_t)nnets + 1, sizeof(int32_t)); out->net_ptr = (int32_t *)calloc((size_t)nnets + 1, sizeof(int32_t));
	out->net_is_global = (uint8_t *)calloc((size_t)nnets + 1, 1);
	count = (int32_t *)calloc((size_t)nnets + 1, sizeof(int32_t)); fill = (int32_t *)calloc((size_t)nnets + 1, sizeof(int32_t));
	if (!out->net_name_ptr || !out->net_ptr || !out->net_is_global || !count || !fill) NL_FAIL(PF_ENOMEM, "out of memory");
	for (int k = 0; k < npins; k++) if (pin_net[k] >= 0) count[pin_net[k]]++;
	{
		size_t chars = 0;
		for (int i = 0; i < nnets; i++) { chars += (size_t)net_names[i].n; out->net_name_ptr[i + 1] = (int32_t)chars; out->net_ptr[i + 1] = out->net_ptr[i] + count[i]; }
		out->net_name_chars = (char *)malloc(chars + 1);
		out->net_block = (int32_t *)malloc(sizeof(int32_t) * (size_t)(out->net_ptr[nnets] > 0 ? out->net_ptr[nnets] : 1));
		out->net_block_pin = (int32_t *)malloc(sizeof(int32_t) * (size_t)(out->net_ptr[nnets] > 0 ? out->net_ptr[nnets] : 1));
		if (!out->net_name_chars || !out->net_block || !out->net_block_pin) NL_FAIL(PF_ENOMEM, "out of memory");
		for (int i = 0; i < nnets; i++) memcpy(out->net_name_chars + out->net_name_ptr[i], net_names[i].s, (size_t)net_names[i].n);
		for (int k = 0; k < out->net_ptr[nnets]; k++) out->net_block[k] = out->net_block_pin[k] = PF_OPEN;
	}
	memset(count, 0xff, sizeof(int32_t) * ((size_t)nnets + 1));      /* now: kind of the net's first receiver pin, -1 = none yet */
	for (int ib = 0; ib < nblocks; ib++)
		for (int k = out->block_pin_ptr[ib]; k < out->block_pin_ptr[ib + 1]; k++) {
			const int net = pin_net[k], pin = k - out->block_pin_ptr[ib];
			if (net < 0) continue;
			const int base = out->net_ptr[net], terms = out->net_ptr[net + 1] - base;
			if (out->block_pin_kind[k] == 1) {
				if (out->net_block[base] != PF_OPEN)
					NL_FAIL(PF_EFORMAT, "%s: net %.*s has two drivers (blocks #%d and #%d)", path, net_names[net].n, net_names[net].s, out->net_block[base], ib);
				out->net_block[base] = ib; out->net_block_pin[base] = pin;
			} else {
				const int glob = out->block_pin_kind[k] == 2;
				if (++fill[net] > terms - 1)      /* read_netlist.c:941-946 */
					NL_FAIL(PF_EFORMAT, "net %.*s #%d inconsistency, expected %d terminals but encountered %d terminals, it is likely net terminal is disconnected in netlist file.",
							net_names[net].n, net_names[net].s, net, terms - 1, fill[net]);
				out->net_block[base + fill[net]] = ib; out->net_block_pin[base + fill[net]] = pin;
				if (count[net] >= 0 && count[net] != glob)      /* read_netlist.c:966-973 */
					NL_FAIL(PF_EFORMAT, "Netlist attempts to connect net %.*s to both global and non-global pins.", net_names[net].n, net_names[net].s);
				count[net] = glob;
				out->net_is_global[net] = (uint8_t)glob;
			}
		}
	for (int i = 0; i < nnets; i++)
		if (out->net_block[out->net_ptr[i]] == PF_OPEN) NL_FAIL(PF_EFORMAT, "%s: net %.*s has no driver", path, net_names[i].n, net_names[i].s);
	for (int k = 0; k < d.nclk; k++) {      /* read_netlist.c:975-979: a circuit clock that reaches a cluster is a global net */
		size_t h = (size_t)fnv1a(d.clk[k].s, (size_t)d.clk[k].n) & tmask;
		while (table[h] >= 0 && !nl_eq(net_names[table[h]], d.clk[k].s, d.clk[k].n)) h = (h + 1) & tmask;
		if (table[h] >= 0 && out->net_ptr[table[h] + 1] - out->net_ptr[table[h]] > 1 && !out->net_is_global[table[h]])
			NL_FAIL(PF_EFORMAT, "%s: circuit clock %.*s drives non-clock pins", path, d.clk[k].n, d.clk[k].s);
	}
	rc = PF_OK;
done:
#undef NL_FAIL
	free(data); free(d.pb); free(d.port); free(d.tok); free(d.clk); free(stack); free(table); free(count); free(fill); free(net_names);
	if (rc != PF_OK) pf_netlist_free(out);
	return rc;
}
