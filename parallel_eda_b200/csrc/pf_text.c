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
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
	for (i = 0; i < count; i++) {
		if (ptr[i + 1] <= ptr[i]) FAIL("%s %d has an empty name", what, i);
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
	for (i = 0; i < n->num_nets; i++) {
		if (n->gpin_ptr[i + 1] < n->gpin_ptr[i]) FAIL("gpin_ptr not monotone at net %d", i);
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

/* ------------------------------------------------------------------ buffered text output */
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
static void ob_flush(outbuf *o) {
	if (o->len && fwrite(o->buf, 1, o->len, o->f) != o->len) o->err = 1;
	o->len = 0;
}
static inline void ob_room(outbuf *o, size_t need) { if (o->len + need > o->cap) ob_flush(o); }
static inline void ob_mem(outbuf *o, const char *s, size_t n) {
	if (n > o->cap) { ob_flush(o); if (fwrite(s, 1, n, o->f) != n) o->err = 1; return; }
	ob_room(o, n);
	memcpy(o->buf + o->len, s, n); o->len += n;
}
#define ob_lit(o, s) ob_mem((o), (s), sizeof(s) - 1)
static inline void ob_int(outbuf *o, long v) {   /* "%d" */
	char t[24];
	int k = 0;
	unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
	ob_room(o, 24);
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

int pf_route_write(const char *path, const pf_problem *p, const pf_names *n, const pf_result *r) {
	outbuf o;
	int inet, k, rc;
	if (!path || !p || !n || !r) return fail(PF_EINVAL, "pf_route_write: NULL argument");
	if (n->num_nets != p->num_nets || r->num_nets != p->num_nets || n->nx != p->nx || n->ny != p->ny)
		return fail(PF_EINVAL, "pf_route_write: problem, names and result disagree on nets or grid");
	if ((rc = ob_open(&o, path)) != 0) return rc;
	ob_lit(&o, "Array size: "); ob_int(&o, p->nx); ob_lit(&o, " x "); ob_int(&o, p->ny); ob_lit(&o, " logic blocks.\n");
	ob_lit(&o, "\nRouting:");
	for (inet = 0; inet < p->num_nets; inet++) {
		const char *name = n->net_name_chars + n->net_name_ptr[inet];
		size_t name_len = (size_t)(n->net_name_ptr[inet + 1] - n->net_name_ptr[inet]);
		ob_lit(&o, "\n\nNet "); ob_int(&o, inet); ob_lit(&o, " ("); ob_mem(&o, name, name_len);
		if (p->net_is_global[inet]) {                                        /* :1394-1412 */
			ob_lit(&o, "): global net connecting:\n\n");
			for (k = n->gpin_ptr[inet]; k < n->gpin_ptr[inet + 1]; k++) {
				int b = n->gpin_block[k];
				ob_lit(&o, "Block ");
				ob_mem(&o, n->block_name_chars + n->block_name_ptr[b], (size_t)(n->block_name_ptr[b + 1] - n->block_name_ptr[b]));
				ob_lit(&o, " (#"); ob_int(&o, b); ob_lit(&o, ") at ("); ob_int(&o, n->block_x[b]); ob_lit(&o, ", ");
				ob_int(&o, n->block_y[b]); ob_lit(&o, "), Pin class "); ob_int(&o, n->gpin_class[k]); ob_lit(&o, ".\n");
			}
			continue;
		}
		ob_lit(&o, ")\n\n");
		if (p->net_ptr[inet + 1] - p->net_ptr[inet] - 1 == 0) {               /* :1337-1339 */
			ob_lit(&o, "\n\nUsed in local cluster only, reserved one CLB pin\n\n");
			continue;
		}
		for (k = r->trace_ptr[inet]; k < r->trace_ptr[inet + 1]; k++) {      /* :1344-1389 */
			int inode = r->trace_node[k], t, ilow, jlow;
			if (inode < 0 || inode >= p->num_nodes) { rc = fail(PF_EINVAL, "net %d: trace node %d out of range", inet, inode); goto out; }
			t = p->type[inode]; ilow = p->xlow[inode]; jlow = p->ylow[inode];
			if (t > PF_CHANY) { rc = fail(PF_EINVAL, "net %d: unexpected traceback element type %d", inet, t); goto out; }
			ob_lit(&o, "Node:\t"); ob_int(&o, inode); ob_lit(&o, "\t"); ob_mem(&o, TYPE_PADDED[t], 6);
			ob_lit(&o, " ("); ob_int(&o, ilow); ob_lit(&o, ","); ob_int(&o, jlow); ob_lit(&o, ") ");
			if (ilow != p->xhigh[inode] || jlow != p->yhigh[inode]) {
				ob_lit(&o, "to ("); ob_int(&o, p->xhigh[inode]); ob_lit(&o, ","); ob_int(&o, p->yhigh[inode]); ob_lit(&o, ") ");
			}
			if (t == PF_CHANX || t == PF_CHANY) ob_lit(&o, " Track: ");
			else if (tile_io(n, ilow, jlow)) ob_lit(&o, " Pad: ");
			else if (t == PF_IPIN || t == PF_OPIN) ob_lit(&o, " Pin: ");
			else ob_lit(&o, " Class: ");
			ob_int(&o, p->ptc_num[inode]); ob_lit(&o, "  \n");
		}
	}
out:
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
	if (*c < '0' || *c > '9') return 0;
	while (*c >= '0' && *c <= '9') { x = x * 10 + (*c - '0'); c++; }
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
		size_t nc = a->cap ? a->cap * 2 : 1 << 16;
		int32_t *nv = (int32_t *)realloc(a->v, nc * sizeof(int32_t));
		if (!nv) return PF_ENOMEM;
		a->v = nv; a->cap = nc;
	}
	a->v[a->n++] = x;
	return 0;
}

int pf_route_read(const char *path, const pf_problem *p, pf_result *r) {
	char *data = NULL;
	size_t len = 0, i;
	const char *s, *end;
	long a, b;
	int line = 0, cur = -1, rc, inet;
	ivec nodes = { 0, 0, 0 };
	int32_t *tptr = NULL;
	int16_t *tsw = NULL;
	memset(r, 0, sizeof(*r));
	g_err[0] = 0;
	if ((rc = slurp(path, &data, &len)) != 0) return rc;
	tptr = (int32_t *)calloc((size_t)p->num_nets + 1, sizeof(int32_t));
	if (!tptr) { rc = PF_ENOMEM; goto done; }
	for (inet = 0; inet <= p->num_nets; inet++) tptr[inet] = -1;
	s = data; end = data + len;
	while (s < end) {
		const char *eol = (const char *)memchr(s, '\n', (size_t)(end - s));
		const char *c = s;
		if (!eol) eol = end;
		line++;
		if (eol == s) { s = eol + 1; continue; }
		if (line == 1) {
			if (!expect(&c, "Array size: ") || !scan_int(&c, &a) || !expect(&c, " x ") || !scan_int(&c, &b)) {
				rc = fail(PF_EFORMAT, "%s:1: not a .route file (no 'Array size:' line)", path); goto done;
			}
			if (a != p->nx || b != p->ny) { rc = fail(PF_EFORMAT, "%s:1: routing of a %ld x %ld array, problem is %d x %d", path, a, b, p->nx, p->ny); goto done; }
		} else if (c[0] == 'N' && c[1] == 'o') {                  /* "Node:\t<id>\t<type> (x,y) [to (x,y) ] <what>: <ptc>  " */
			long id, x, y, xh, yh, ptc;
			int t;
			if (cur < 0 || p->net_is_global[cur]) { rc = fail(PF_EFORMAT, "%s:%d: Node line outside a routed net", path, line); goto done; }
			if (!expect(&c, "Node:\t") || !scan_int(&c, &id) || *c != '\t') { rc = fail(PF_EFORMAT, "%s:%d: malformed Node line", path, line); goto done; }
			c++;
			while (*c == ' ') c++;
			for (t = 0; t < 6; t++) { size_t tl = strlen(TYPE_NAME[t]); if (strncmp(c, TYPE_NAME[t], tl) == 0 && c[tl] == ' ') { c += tl; break; } }
			if (t == 6 || !expect(&c, " (") || !scan_int(&c, &x) || !expect(&c, ",") || !scan_int(&c, &y) || !expect(&c, ") ")) {
				rc = fail(PF_EFORMAT, "%s:%d: malformed Node line", path, line); goto done;
			}
			xh = x; yh = y;
			if (c[0] == 't') {
				if (!expect(&c, "to (") || !scan_int(&c, &xh) || !expect(&c, ",") || !scan_int(&c, &yh) || !expect(&c, ") ")) {
					rc = fail(PF_EFORMAT, "%s:%d: malformed 'to (x,y)'", path, line); goto done;
				}
			}
			while (c < eol && *c != ':') c++;                        /* " Pad" / " Pin" / " Track" / " Class" */
			if (c >= eol || c[1] != ' ') { rc = fail(PF_EFORMAT, "%s:%d: malformed Node line", path, line); goto done; }
			c += 2;
			if (!scan_int(&c, &ptc)) { rc = fail(PF_EFORMAT, "%s:%d: no track / pin / class number", path, line); goto done; }
			if (id < 0 || id >= p->num_nodes) { rc = fail(PF_EFORMAT, "%s:%d: rr node %ld out of range", path, line, id); goto done; }
			if (p->type[id] != t || p->xlow[id] != x || p->ylow[id] != y || p->xhigh[id] != xh || p->yhigh[id] != yh || p->ptc_num[id] != ptc) {
				rc = fail(PF_EFORMAT, "%s:%d: rr node %ld is %s (%d,%d)-(%d,%d) ptc %d in the problem: the file belongs to another rr graph",
						path, line, id, TYPE_NAME[p->type[id] <= PF_CHANY ? p->type[id] : 0], p->xlow[id], p->ylow[id], p->xhigh[id], p->yhigh[id], p->ptc_num[id]);
				goto done;
			}
			if ((rc = iv_push(&nodes, (int32_t)id)) != 0) goto done;
		} else if (c[0] == 'N' && c[1] == 'e') {                  /* "Net <i> (<name>)" or "...): global net connecting:" */
			if (!expect(&c, "Net ") || !scan_int(&c, &a) || !expect(&c, " (")) { rc = fail(PF_EFORMAT, "%s:%d: malformed Net line", path, line); goto done; }
			if (a != cur + 1 || a >= p->num_nets) { rc = fail(PF_EFORMAT, "%s:%d: net %ld follows net %d (problem has %d nets)", path, line, a, cur, p->num_nets); goto done; }
			cur = (int)a;
			tptr[cur] = (int32_t)nodes.n;
			{
				static const char GLOB[] = "): global net connecting:";      /* matched at the tail: names may contain ')' */
				const size_t gl = sizeof(GLOB) - 1;
				int glob = (size_t)(eol - c) >= gl && memcmp(eol - gl, GLOB, gl) == 0;
				if (!glob && eol[-1] != ')') { rc = fail(PF_EFORMAT, "%s:%d: malformed Net line", path, line); goto done; }
				if (glob != (p->net_is_global[cur] != 0)) { rc = fail(PF_EFORMAT, "%s:%d: net %d global in one of file / problem only", path, line, cur); goto done; }
			}
		} else if (strncmp(c, "Routing:", 8) == 0 || strncmp(c, "Block ", 6) == 0 || strncmp(c, "Used in local", 13) == 0) {
			/* nothing to take from these */
		} else {
			rc = fail(PF_EFORMAT, "%s:%d: unrecognised line", path, line); goto done;
		}
		s = eol + 1;
	}
	if (cur != p->num_nets - 1) { rc = fail(PF_EFORMAT, "%s: %d nets in the file, %d in the problem", path, cur + 1, p->num_nets); goto done; }
	tptr[p->num_nets] = (int32_t)nodes.n;
	tsw = (int16_t *)malloc(sizeof(int16_t) * (nodes.n ? nodes.n : 1));
	if (!tsw) { rc = PF_ENOMEM; goto done; }
	{
		long wl = 0;
		int serial = 0;
		for (inet = 0; inet < p->num_nets; inet++) {
			size_t lo = (size_t)tptr[inet], hi = (size_t)tptr[inet + 1];
			for (i = lo; i < hi; i++) {
				int u = nodes.v[i], e, found = -1;
				serial += (inet + 1) * (p->xlow[u] * (p->nx + 1) - p->yhigh[u]);     /* get_serial_num, route_common.c:224-254 */
				serial -= p->ptc_num[u] * (inet + 1) * 10;
				serial -= p->type[u] * (inet + 1) * 100;
				serial %= 2000000000;
				if (p->type[u] == PF_SINK) { tsw[i] = PF_OPEN; continue; }
				if (i + 1 >= hi) { rc = fail(PF_EFORMAT, "%s: net %d does not end at a SINK", path, inet); goto done; }
				for (e = p->row_ptr[u]; e < p->row_ptr[u + 1]; e++)
					if (p->edge_to[e] == nodes.v[i + 1]) { found = e; break; }
				if (found < 0) { rc = fail(PF_EFORMAT, "%s: net %d: no rr edge %d -> %d", path, inet, u, nodes.v[i + 1]); goto done; }
				tsw[i] = p->edge_sw[found];
			}
			/* get_num_bends_and_length, base/stats.c:355-409: the first element and every element after a SINK are skipped */
			for (i = lo + 1; i < hi; i++) {
				int u = nodes.v[i], t = p->type[u];
				if (t == PF_SINK) { i++; continue; }
				if (t == PF_CHANX || t == PF_CHANY) wl += 1 + p->xhigh[u] - p->xlow[u] + p->yhigh[u] - p->ylow[u];
			}
		}
		r->total_wirelength = (int32_t)wl;
		r->serial_num = serial;
	}
	r->num_nets = p->num_nets;
	r->trace_ptr = tptr; tptr = NULL;
	r->trace_node = nodes.v ? nodes.v : (int32_t *)calloc(1, sizeof(int32_t)); nodes.v = NULL;
	r->trace_switch = tsw; tsw = NULL;
	rc = PF_OK;
done:
	free(data); free(tptr); free(tsw); free(nodes.v);
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
