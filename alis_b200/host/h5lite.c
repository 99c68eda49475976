/*
 * h5lite.c -- the small part of HDF5 that ALiS uses, without libhdf5: one float dataset of rank
 * 1..3 per file (reference world.c:469-537: h5write creates "data" in the root group with default
 * properties, h5read opens the first object of the root group).
 *
 * Writer: the classic on-disk layout libhdf5 produces for exactly that call sequence (HDF5 File
 * Format Specification, version-0 superblock): superblock, root group object header with a symbol
 * table message, group B-tree node, local heap, one symbol table node, the dataset's version-1
 * object header (dataspace, IEEE little-endian float32 datatype, fill value, contiguous layout)
 * and the raw data.  Reader: walks the same structures (superblock v0/v1, B-tree to its leftmost
 * leaf, first symbol, object header incl. continuation blocks) and accepts contiguous or compact
 * float32/float64 datasets of either byte order; chunked or filtered datasets and new-style
 * (superblock v2+) files are reported as unsupported.
 *
 * libhdf5 is not available in the build image: the writer is validated by this reader and by the
 * byte-level checks in tests/test_h5lite.py, not against the library itself.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "h5lite.h"

#define UNDEF 0xFFFFFFFFFFFFFFFFull
static const unsigned char SIGNATURE[8] = { 0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n' };
enum { LEAF_K = 4, INTERNAL_K = 16 };

/* ---- little-endian byte buffer --------------------------------------------------------------- */
typedef struct { unsigned char *p; size_t n, cap; } buf;
static void put(buf *b, const void *src, size_t n)
{
	if (b->n + n > b->cap) { b->cap = 2*(b->n + n) + 256; b->p = realloc(b->p, b->cap); }
	if (src) memcpy(b->p + b->n, src, n); else memset(b->p + b->n, 0, n);
	b->n += n;
}
static void put_le(buf *b, uint64_t v, int bytes) { unsigned char t[8]; for (int q=0; q<bytes; q++) t[q] = (unsigned char)(v >> (8*q)); put(b, t, bytes); }
static void pad_to(buf *b, size_t n) { if (b->n < n) put(b, 0, n - b->n); }

/* one object-header message (version-1 header): type, size, flags, 3 reserved, data padded to 8 */
static void message(buf *b, int type, const buf *data)
{
	const size_t padded = (data->n + 7) & ~(size_t)7;
	put_le(b, type, 2); put_le(b, padded, 2); put_le(b, 0, 1); put_le(b, 0, 3);
	put(b, data->p, data->n);
	put(b, 0, padded - data->n);
}

static void object_header(buf *b, int nmessages, const buf *messages)
{
	put_le(b, 1, 1); put_le(b, 0, 1); put_le(b, nmessages, 2);      /* version, reserved, message count */
	put_le(b, 1, 4);                                                /* object reference count           */
	put_le(b, messages->n, 4);                                      /* bytes of message data            */
	put_le(b, 0, 4);                                                /* alignment of the prefix to 8     */
	put(b, messages->p, messages->n);
}

int h5lite_write(const char *path, int rank, const unsigned long long *dims, const float *data)
{
	if (rank < 1 || rank > 3) return -1;
	size_t count = 1;
	for (int d=0; d<rank; d++) count *= dims[d];

	/* fixed addresses of the metadata blocks */
	const uint64_t root_oh = 96, btree = root_oh + 16 + 24, heap = btree + 24 + (2*INTERNAL_K+1)*8 + 2*INTERNAL_K*8;
	const uint64_t heap_data = heap + 32, heap_size = 32, snod = heap_data + heap_size, dset_oh = snod + 8 + 2*LEAF_K*40;
	const uint64_t dset_msgs = 96 + 8*(uint64_t)rank, raw = dset_oh + 16 + dset_msgs, eof = raw + 4*(uint64_t)count;

	buf f = { 0, 0, 0 };
	/* superblock, version 0 */
	put(&f, SIGNATURE, 8);
	put_le(&f, 0, 1); put_le(&f, 0, 1); put_le(&f, 0, 1); put_le(&f, 0, 1);   /* superblock, free space, root entry versions; reserved */
	put_le(&f, 0, 1); put_le(&f, 8, 1); put_le(&f, 8, 1); put_le(&f, 0, 1);   /* shared header version; offset size; length size; reserved */
	put_le(&f, LEAF_K, 2); put_le(&f, INTERNAL_K, 2); put_le(&f, 0, 4);       /* group leaf K, internal K, consistency flags */
	put_le(&f, 0, 8); put_le(&f, UNDEF, 8); put_le(&f, eof, 8); put_le(&f, UNDEF, 8);   /* base, free space, end of file, driver info */
	put_le(&f, 0, 8); put_le(&f, root_oh, 8); put_le(&f, 1, 4); put_le(&f, 0, 4);       /* root entry: name offset, header, cache type 1 */
	put_le(&f, btree, 8); put_le(&f, heap, 8);                                          /*   scratch pad: B-tree, heap */

	/* root group object header: symbol table message */
	{
		buf m = { 0, 0, 0 }, st = { 0, 0, 0 };
		put_le(&st, btree, 8); put_le(&st, heap, 8);
		message(&m, 0x0011, &st);
		pad_to(&f, root_oh);
		object_header(&f, 1, &m);
		free(m.p); free(st.p);
	}
	/* group B-tree: one leaf-level node with one child */
	pad_to(&f, btree);
	put(&f, "TREE", 4); put_le(&f, 0, 1); put_le(&f, 0, 1); put_le(&f, 1, 2);           /* node type 0 (group), level 0, 1 entry */
	put_le(&f, UNDEF, 8); put_le(&f, UNDEF, 8);                                         /* siblings */
	put_le(&f, 0, 8); put_le(&f, snod, 8); put_le(&f, 8, 8);                            /* key 0 (""), child 0, key 1 ("data") */
	pad_to(&f, heap);
	/* local heap: "" at 0, "data" at 8, one free block behind */
	put(&f, "HEAP", 4); put_le(&f, 0, 1); put_le(&f, 0, 3);
	put_le(&f, heap_size, 8); put_le(&f, 16, 8); put_le(&f, heap_data, 8);              /* segment size, free-list head, segment address */
	put_le(&f, 0, 8);
	put(&f, "data\0\0\0", 8);
	put_le(&f, 1, 8); put_le(&f, heap_size - 16, 8);                                    /* free block: no next (1), its size */
	/* symbol table node with the dataset's entry */
	pad_to(&f, snod);
	put(&f, "SNOD", 4); put_le(&f, 1, 1); put_le(&f, 0, 1); put_le(&f, 1, 2);
	put_le(&f, 8, 8); put_le(&f, dset_oh, 8); put_le(&f, 0, 4); put_le(&f, 0, 4); put_le(&f, 0, 8); put_le(&f, 0, 8);
	pad_to(&f, dset_oh);
	/* dataset object header */
	{
		buf m = { 0, 0, 0 }, d = { 0, 0, 0 };
		put_le(&d, 1, 1); put_le(&d, rank, 1); put_le(&d, 0, 1); put_le(&d, 0, 1); put_le(&d, 0, 4);     /* dataspace v1, no maximum sizes */
		for (int q=0; q<rank; q++) put_le(&d, dims[q], 8);
		message(&m, 0x0001, &d); d.n = 0;
		put_le(&d, 0x11, 1); put_le(&d, 0x20, 1); put_le(&d, 31, 1); put_le(&d, 0, 1); put_le(&d, 4, 4); /* float v1, little endian, implied mantissa msb, sign bit 31, 4 bytes */
		put_le(&d, 0, 2); put_le(&d, 32, 2); put_le(&d, 23, 1); put_le(&d, 8, 1); put_le(&d, 0, 1); put_le(&d, 23, 1); put_le(&d, 127, 4);
		message(&m, 0x0003, &d); d.n = 0;
		put_le(&d, 2, 1); put_le(&d, 2, 1); put_le(&d, 2, 1); put_le(&d, 1, 1); put_le(&d, 0, 4);         /* fill value v2: late allocation, write if set, default value */
		message(&m, 0x0005, &d); d.n = 0;
		put_le(&d, 3, 1); put_le(&d, 1, 1); put_le(&d, raw, 8); put_le(&d, 4*(uint64_t)count, 8);         /* layout v3, contiguous */
		message(&m, 0x0008, &d);
		if (m.n != dset_msgs) { free(m.p); free(d.p); free(f.p); return -2; }
		object_header(&f, 4, &m);
		free(m.p); free(d.p);
	}
	if (f.n != raw) { free(f.p); return -2; }

	FILE *F = fopen(path, "wb");
	if (!F) { free(f.p); return -3; }
	int ok = fwrite(f.p, 1, f.n, F) == f.n;
	{	/* raw data, little endian */
		const uint16_t probe = 1;
		if (*(const unsigned char*)&probe) ok = ok && fwrite(data, 4, count, F) == count;
		else for (size_t q=0; q<count && ok; q++) { uint32_t v; unsigned char t[4]; memcpy(&v, data+q, 4); for (int s=0; s<4; s++) t[s] = (unsigned char)(v >> (8*s)); ok = fwrite(t, 1, 4, F) == 4; }
	}
	ok = fclose(F) == 0 && ok;
	free(f.p);
	return ok ? 0 : -3;
}

/* ---- reader ------------------------------------------------------------------------------------ */
typedef struct { FILE *F; uint64_t base; int so, sl; char err[160]; } rd;

static int rd_at(rd *r, uint64_t addr, void *dst, size_t n)
{
	if (addr == UNDEF || fseek(r->F, (long)(r->base + addr), SEEK_SET) || fread(dst, 1, n, r->F) != n) { snprintf(r->err, sizeof r->err, "short read at %llu", (unsigned long long)addr); return -1; }
	return 0;
}
static uint64_t le(const unsigned char *p, int bytes) { uint64_t v = 0; for (int q=0; q<bytes; q++) v |= (uint64_t)p[q] << (8*q); return v; }

typedef struct { int rank; unsigned long long dims[3]; int elem, is_float, big_endian; int layout; uint64_t addr, size; unsigned char *compact; } dinfo;

/* every field of a message is read only after its extent has been checked against the message size the
 * file declares (files come from users) */
#define NEED_BYTES(n) do { if ((uint64_t)(n) > size) { snprintf(r->err, sizeof r->err, "truncated header message (type 0x%04x)", type); free(m); return -1; } } while (0)
static int parse_messages(rd *r, uint64_t addr, uint64_t bytes, int *left, dinfo *D, int depth)
{
	if (bytes > (1u << 26)) { snprintf(r->err, sizeof r->err, "implausible object header size"); return -1; }
	unsigned char *m = malloc(bytes ? bytes : 1);
	if (!m) { snprintf(r->err, sizeof r->err, "out of memory"); return -1; }
	if (rd_at(r, addr, m, bytes)) { free(m); return -1; }
	uint64_t pos = 0;
	while (*left > 0 && pos + 8 <= bytes) {
		const int type = (int)le(m+pos, 2); const uint64_t size = le(m+pos+2, 2);
		const unsigned char *d = m + pos + 8;
		if (pos + 8 + size > bytes) break;
		(*left)--;
		if (type == 0x0001) {                                   /* dataspace v1 / v2 */
			NEED_BYTES(2);
			const int ver = d[0]; D->rank = d[1];
			const size_t lead = ver == 1 ? 8 : 4;
			const unsigned char *dim = d + lead;
			if (D->rank < 1 || D->rank > 3) { snprintf(r->err, sizeof r->err, "dataset rank %d (1..3 supported)", D->rank); free(m); return -1; }
			NEED_BYTES(lead + (size_t)D->rank*r->sl);
			for (int q=0; q<D->rank; q++) D->dims[q] = le(dim + q*r->sl, r->sl);
		}
		else if (type == 0x0003) {                              /* datatype */
			NEED_BYTES(8);
			const int cls = d[0] & 15;
			D->is_float = cls == 1; D->big_endian = d[1] & 1; D->elem = (int)le(d+4, 4);
		}
		else if (type == 0x0008) {                              /* data layout */
			NEED_BYTES(2);
			if (d[0] != 3) { snprintf(r->err, sizeof r->err, "data layout message version %d (3 supported)", d[0]); free(m); return -1; }
			D->layout = d[1];
			if (d[1] == 1) { NEED_BYTES(2 + r->so + r->sl); D->addr = le(d+2, r->so); D->size = le(d+2+r->so, r->sl); }
			else if (d[1] == 0) {
				NEED_BYTES(4);
				D->size = le(d+2, 2);
				NEED_BYTES(4 + D->size);                            /* the data must lie inside the message */
				free(D->compact);                                   /* a second layout message replaces the first */
				D->compact = malloc(D->size ? D->size : 1);
				if (!D->compact) { snprintf(r->err, sizeof r->err, "out of memory"); free(m); return -1; }
				memcpy(D->compact, d+4, D->size);
			}
		}
		else if (type == 0x000B) { snprintf(r->err, sizeof r->err, "filtered (compressed) datasets are not supported"); free(m); return -1; }
		else if (type == 0x0010 && depth < 8) {                 /* continuation block */
			NEED_BYTES(r->so + r->sl);
			if (parse_messages(r, le(d, r->so), le(d+r->so, r->sl), left, D, depth+1)) { free(m); return -1; }
		}
		pos += 8 + size;
	}
	free(m);
	return 0;
}

void *h5lite_read(const char *path, int *rank_out, unsigned long long dims_out[3], char *err, size_t errlen)
{
	rd r; memset(&r, 0, sizeof r);
	dinfo D; memset(&D, 0, sizeof D);
	float *out = 0;
	#define FAIL(...) do { snprintf(r.err, sizeof r.err, __VA_ARGS__); goto done; } while (0)
	r.F = fopen(path, "rb");
	if (!r.F) { snprintf(err, errlen, "cannot open %s", path); return 0; }
	unsigned char sb[128];
	/* the superblock may sit at 0, 512, 1024, ... (a user block in front) */
	for (r.base = 0; ; r.base = r.base ? 2*r.base : 512) {
		if (rd_at(&r, 0, sb, 8)) FAIL("%s is not an HDF5 file", path);
		if (!memcmp(sb, SIGNATURE, 8)) break;
		if (r.base > (1u<<24)) FAIL("%s is not an HDF5 file", path);
	}
	if (rd_at(&r, 0, sb, 24)) goto done;
	const int sbver = sb[8];
	if (sbver > 1) FAIL("superblock version %d (new-style file): only the classic layout (versions 0 and 1) is supported", sbver);
	r.so = sb[13]; r.sl = sb[14];
	if (r.so != 8 || r.sl != 8) FAIL("offset/length size %d/%d (8/8 supported)", r.so, r.sl);
	const int lead = 24 + (sbver == 1 ? 4 : 0);                   /* v1 adds indexed-storage K + 2 reserved bytes */
	if (rd_at(&r, 0, sb, lead + 4*8 + 40)) goto done;
	const unsigned char *root = sb + lead + 4*8;
	uint64_t node = le(root + 24, 8);                             /* scratch pad: B-tree address */
	if (le(root + 16, 4) != 1) {                                  /* not cached: read the symbol table message of the root header */
		FAIL("root group without cached symbol table");
	}
	/* descend to the leftmost leaf of the group B-tree */
	uint64_t snod = UNDEF;
	for (int guard=0; guard<32; guard++) {
		unsigned char t[24 + 3*8];
		if (rd_at(&r, node, t, sizeof t)) goto done;
		if (memcmp(t, "TREE", 4) || t[4] != 0) FAIL("bad group B-tree node");
		if (!le(t+6, 2)) FAIL("the root group is empty");
		const uint64_t child = le(t + 24 + 8, 8);
		if (t[5] == 0) { snod = child; break; }
		node = child;
	}
	if (snod == UNDEF) FAIL("group B-tree too deep");
	unsigned char s[8 + 40];
	if (rd_at(&r, snod, s, sizeof s)) goto done;
	if (memcmp(s, "SNOD", 4) || !le(s+6, 2)) FAIL("bad symbol table node");
	const uint64_t oh = le(s + 8 + 8, 8);
	unsigned char h[16];
	if (rd_at(&r, oh, h, 16)) goto done;
	if (h[0] != 1) FAIL("object header version %d (1 supported)", h[0]);
	int left = (int)le(h+2, 2);
	if (parse_messages(&r, oh + 16, le(h+8, 4), &left, &D, 0)) goto done;
	if (!D.rank || !D.elem) FAIL("the first object of the file is not a simple dataset");
	if (!D.is_float || (D.elem != 4 && D.elem != 8)) FAIL("dataset type is not a 4- or 8-byte float");
	if (D.layout != 0 && D.layout != 1) FAIL("chunked datasets are not supported");
	{
		size_t count = 1;
		for (int q=0; q<D.rank; q++) {                                /* overflow-safe product, bounded by what a float*** of the library can hold */
			if (D.dims[q] && count > ((size_t)1 << 40)/D.dims[q]) FAIL("dataset too large");
			count *= D.dims[q];
		}
		unsigned char *raw = D.compact;
		if (D.layout == 1) {
			raw = malloc(count*D.elem ? count*D.elem : 1);
			if (!raw) FAIL("out of memory (%zu elements)", count);
			if (D.addr == UNDEF) memset(raw, 0, count*D.elem);     /* never written: fill value */
			else if (rd_at(&r, D.addr, raw, count*D.elem)) { free(raw); goto done; }
		}
		else if (!raw || D.size < count*D.elem) FAIL("compact dataset shorter than its dataspace");
		out = malloc(sizeof(float)*(count ? count : 1));
		if (!out) { if (D.layout == 1) free(raw); FAIL("out of memory (%zu elements)", count); }
		for (size_t q=0; q<count; q++) {
			unsigned char t[8];
			for (int b=0; b<D.elem; b++) t[b] = raw[q*D.elem + (D.big_endian ? D.elem-1-b : b)];
			if (D.elem == 4) { uint32_t v = (uint32_t)le(t, 4); memcpy(out+q, &v, 4); }
			else { uint64_t v = le(t, 8); double dv; memcpy(&dv, &v, 8); out[q] = (float)dv; }
		}
		if (D.layout == 1) free(raw);
		*rank_out = D.rank;
		for (int q=0; q<3; q++) dims_out[q] = q < D.rank ? D.dims[q] : 0;
	}
done:
	#undef FAIL
	if (!out && err) snprintf(err, errlen, "%s", r.err[0] ? r.err : "read error");
	free(D.compact);
	fclose(r.F);
	return out;
}
