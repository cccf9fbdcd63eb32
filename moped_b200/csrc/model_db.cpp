// model_db.cpp — model-database loader: `.moped.xml` -> packed row tables -> device.
//
// Replaces, for the hot path's input side, the reference's model load:
//   sXML::process                       moped2/libmoped/include/sXML.hpp:55-127   (the XML subset reader)
//   MopedPimpl::addModel(sXML&)         moped2/libmoped/src/moped.cpp:100-137     (Points -> Model::IPs[desc_type])
//   MopedPimpl::addModel(SP_Model&)     moped2/libmoped/src/moped.cpp:138-149     (same name replaces in place)
//   MATCH_ANN_CPU::Update               moped2/libmoped/src/match/MATCH_ANN_CPU.hpp:72-109 (row order, norm())
// The file format stays byte-compatible with what the reference's tools write
// (moped-modeling-py/src/MopedModeling.py:900-1000, moped3d/modeling/sfm_export_xml.m): a root element with a
// `name` property, a `Points` child, one child per 3-D point carrying `p3d`, `desc_type` and `desc` properties
// (space separated decimals); `Observation` grandchildren, `Cameras`, `Openrave` and comments are skipped.
//
// Design: no DOM. One forward pass over the memory-mapped text; the only strings kept are the model name and the
// descriptor-type keys; numbers go straight into flat float arrays. Files are independent, so a list of files is
// parsed by a pool of host threads and appended in list order. `pack` lays the rows out exactly as the reference's
// matcher numbers them (models in list order, points in file order) and normalises each descriptor with the stage
// class's expression; `save`/`load` keep that packed form as a binary cache so that a 1 M-descriptor database
// (about 1 GB of XML text) is parsed once.
//
// Lexical rules follow sXML: a name/property token ends at whitespace, '>' or '='; a property value is the text
// between the next two '"', where a backslash drops itself and `\n` becomes a newline; `<!-- ... -->` comments may
// precede an element; an element whose name starts with '/' closes its parent. Numbers are read the way
// `istream >> float` reads them (optional sign, digits, '.', digits, exponent), reading stops at the first token
// that is not a number (moped.cpp:124-130), and conversion is strtof (correctly rounded, like libstdc++'s num_get).
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/moped_cuda.h"

namespace {

struct PointSet {                       // Model::IPs[desc_type] (moped.hpp:205-222), flattened
	std::vector<float> xyz;             // 3 per point
	std::vector<float> desc;            // concatenated descriptors
	std::vector<uint32_t> desc_off;     // n_points + 1 offsets into desc (descriptor lengths may differ in a broken file)
	PointSet() { desc_off.push_back(0); }
	size_t n() const { return desc_off.size() - 1; }
};

struct ModelRec {
	std::string name;
	float bbox[6];
	std::map<std::string, PointSet> ips;    // std::map like the reference: iteration order = key order
	bool has_points;
	ModelRec() : has_points(false) {
		for (int i = 0; i < 3; i++) { bbox[i] = 10E10f; bbox[3 + i] = -10E10f; }
	}
};

struct Packed {                         // rows of one descriptor type in matcher order
	std::string desc_type;
	int D = 0;
	std::vector<float> desc, xyz;
	std::vector<int32_t> model_of_row, n_pts;
	bool normalised = false;
};

// ---------------------------------------------------------------------------------------------
// sXML-compatible forward scanner
// ---------------------------------------------------------------------------------------------
struct Cursor {
	const char *p, *end;
	bool eof() const { return p >= end; }
	int peek() const { return p < end ? (unsigned char)*p : -1; }
};

struct ParseError { std::string what; };

inline bool is_space(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// sXML::getToken: characters up to whitespace, '>' or '='; then trailing whitespace is consumed
inline void get_token(Cursor &c, const char *&b, const char *&e) {
	b = c.p;
	while (!c.eof() && !is_space(*c.p) && *c.p != '>' && *c.p != '=') c.p++;
	e = c.p;
	if (c.eof()) throw ParseError{"unexpected end of file inside a tag"};
	while (!c.eof() && is_space(*c.p)) c.p++;
}

inline void skip_to_lt(Cursor &c) {
	const char *q = (const char *)memchr(c.p, '<', (size_t)(c.end - c.p));
	if (!q) throw ParseError{"unexpected end of file (no further element)"};
	c.p = q + 1;
}

// Reads "<name", skipping comments. Returns false if the element is a closing/empty tag (name empty, starts with
// '/', or ends with '/'), like the early returns of sXML::process.
inline bool open_tag(Cursor &c, const char *&nb, const char *&ne) {
	skip_to_lt(c);
	get_token(c, nb, ne);
	while (ne - nb == 3 && memcmp(nb, "!--", 3) == 0) {
		int state = 0;
		for (;;) {
			if (c.eof()) throw ParseError{"unterminated comment"};
			if (state >= 2 && *c.p == '>') break;
			state = (*c.p == '-') ? state + 1 : 0;
			c.p++;
		}
		skip_to_lt(c);
		get_token(c, nb, ne);
	}
	return !(nb == ne || *nb == '/' || ne[-1] == '/');
}

// One `name="value"` pair; returns false when the property list has ended. The raw value range is returned
// (escapes are resolved by unescape() only for the few values that are kept as strings).
inline bool next_property(Cursor &c, const char *&kb, const char *&ke, const char *&vb, const char *&ve) {
	if (c.eof()) throw ParseError{"unexpected end of file in a property list"};
	if (*c.p == '/') return false;
	get_token(c, kb, ke);
	if (kb == ke || c.eof() || *c.p != '=') return false;
	const char *q = (const char *)memchr(c.p, '"', (size_t)(c.end - c.p));
	if (!q) throw ParseError{"property value without opening quote"};
	vb = q + 1;
	const char *r = vb;
	for (;;) {
		if (r >= c.end) throw ParseError{"unterminated property value"};
		if (*r == '"') break;
		if (*r == '\\') { r++; if (r < c.end && *r == 'n') r++; if (r >= c.end) throw ParseError{"unterminated property value"}; }
		r++;
	}
	ve = r;
	c.p = r + 1;
	while (!c.eof() && is_space(*c.p)) c.p++;
	return true;
}

std::string unescape(const char *b, const char *e) {
	std::string s;
	for (const char *r = b; r < e;) {
		if (*r == '\\') { r++; if (r < e && *r == 'n') { s += '\n'; r++; } if (r >= e) break; }
		s += *r++;
	}
	return s;
}

inline bool key_is(const char *b, const char *e, const char *lit) {
	size_t n = strlen(lit);
	return (size_t)(e - b) == n && memcmp(b, lit, n) == 0;
}

// Correctly rounded decimal -> float without strtof for the common case (Clinger's fast path): at most 15 significant
// digits and |exponent| <= 22 make mantissa and power of ten exact doubles, so one multiply or divide gives the
// correctly rounded DOUBLE; narrowing it to float is also correct unless that double sits exactly on the midpoint of
// two floats (the decimal may lie on either side of it) or the result leaves the normal float range — those cases, and
// longer inputs, return false and go through strtof. Token grammar already checked by the caller.
inline bool parse_float_fast(const char *p, const char *e, float &out) {
	static const double p10[23] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19,
	                                1e20, 1e21, 1e22 };
	bool neg = false;
	if (p < e && (*p == '+' || *p == '-')) { neg = *p == '-'; p++; }
	uint64_t mant = 0;
	int digits = 0, exp10 = 0;
	bool seen_nonzero = false;
	for (; p < e && *p >= '0' && *p <= '9'; p++) {
		if (*p != '0') seen_nonzero = true;
		if (seen_nonzero) { if (++digits > 15) return false; mant = mant * 10 + (uint64_t)(*p - '0'); }
	}
	if (p < e && *p == '.') {
		for (p++; p < e && *p >= '0' && *p <= '9'; p++) {
			if (*p != '0') seen_nonzero = true;
			if (seen_nonzero) { if (++digits > 15) return false; mant = mant * 10 + (uint64_t)(*p - '0'); }
			exp10--;
		}
	}
	if (p < e && (*p == 'e' || *p == 'E')) {
		p++;
		bool eneg = false;
		if (p < e && (*p == '+' || *p == '-')) { eneg = *p == '-'; p++; }
		if (p >= e) return false;
		int x = 0;
		for (; p < e && *p >= '0' && *p <= '9'; p++) { x = x * 10 + (*p - '0'); if (x > 10000) return false; }
		exp10 += eneg ? -x : x;
	}
	if (p != e) return false;
	if (mant == 0) { out = neg ? -0.0f : 0.0f; return true; }
	if (exp10 < -22 || exp10 > 22) return false;
	double d = (double)mant;
	d = exp10 < 0 ? d / p10[-exp10] : d * p10[exp10];
	if (!(d >= 1.17549435e-38 && d <= 3.4028234e+38)) return false;      // subnormal or overflowing floats: let strtof decide
	uint64_t bits;
	memcpy(&bits, &d, 8);
	if ((bits & 0x1FFFFFFFull) == 0x10000000ull) return false;             // exactly between two floats
	out = (float)(neg ? -d : d);
	return true;
}

// `istream >> float` until it fails: appends the numbers of [b, e) to out; returns how many were read.
// A value range containing a backslash is unescaped first (never the case in files the tools write).
size_t read_floats(const char *b, const char *e, std::vector<float> &out, size_t max_count) {
	std::string tmp;
	if (memchr(b, '\\', (size_t)(e - b))) { tmp = unescape(b, e); b = tmp.data(); e = b + tmp.size(); }
	size_t n = 0;
	char buf[64];
	const char *p = b;
	while (n < max_count) {
		while (p < e && is_space(*p)) p++;
		if (p >= e) break;
		// libstdc++ num_get::_M_extract_float grammar: [+-] digits [. digits] [eE [+-] digits]
		const char *t = p;
		if (t < e && (*t == '+' || *t == '-')) t++;
		const char *d0 = t;
		while (t < e && *t >= '0' && *t <= '9') t++;
		size_t nd = (size_t)(t - d0);
		if (t < e && *t == '.') { t++; const char *f0 = t; while (t < e && *t >= '0' && *t <= '9') t++; nd += (size_t)(t - f0); }
		if (nd == 0) break;                                 // not a number: extraction fails, the loop ends
		if (t < e && (*t == 'e' || *t == 'E')) {
			const char *x = t + 1;
			if (x < e && (*x == '+' || *x == '-')) x++;
			const char *x0 = x;
			while (x < e && *x >= '0' && *x <= '9') x++;
			if (x > x0) t = x;
			else { t = x; }                                 // "1e" / "1e+": the stream consumes it and the conversion fails
		}
		size_t len = (size_t)(t - p);
		float v;
		if (parse_float_fast(p, t, v)) {
		} else if (len < sizeof buf) {
			memcpy(buf, p, len); buf[len] = 0;
			char *endp = nullptr;
			errno = 0;
			v = strtof(buf, &endp);
			if (endp != buf + len) break;                   // dangling exponent: failbit
			if (std::isinf(v)) break;                       // overflow: num_get sets failbit, the value is not kept
		} else {
			std::string big(p, len);
			char *endp = nullptr;
			v = strtof(big.c_str(), &endp);
			if (endp != big.c_str() + len || std::isinf(v)) break;
		}
		out.push_back(v);
		n++;
		p = t;
	}
	return n;
}

// Skips the rest of the current element (cursor just after its property list): children and the closing tag,
// leaving the cursor where sXML::process leaves it (before the '>' that ends the element).
void skip_children(Cursor &c);

// returns true when the child was a closing tag (name starts with '/')
inline bool is_closing(const char *nb, const char *ne) { return nb != ne && *nb == '/'; }

void finish_element(Cursor &c) {                 // sXML: `while( in.peek() != '>' ) in.get();`
	const char *q = (const char *)memchr(c.p, '>', (size_t)(c.end - c.p));
	if (!q) throw ParseError{"unexpected end of file (unterminated tag)"};
	c.p = q;
}

void skip_children(Cursor &c) {
	while (!c.eof() && *c.p == '>') {
		const char *nb, *ne;
		Cursor save = c;
		if (!open_tag(c, nb, ne)) {
			if (is_closing(nb, ne)) return;          // parent returns immediately, cursor stays after the token
			(void)save;
			continue;                                // empty-named or "x/" element: ignored, loop re-tests peek
		}
		const char *kb, *ke, *vb, *ve;
		while (next_property(c, kb, ke, vb, ve)) {}
		skip_children(c);
		finish_element(c);
	}
	finish_element(c);
}

// children of <Points>: every named child is a point (moped.cpp:117-133)
void parse_points(Cursor &c, ModelRec &m) {
	std::vector<float> p3;
	std::string type_key;
	PointSet *cur = nullptr;
	std::string cur_key;
	bool have_cur = false;
	while (!c.eof() && *c.p == '>') {
		const char *nb, *ne;
		const bool opened = open_tag(c, nb, ne);
		if (!opened) {
			if (is_closing(nb, ne)) return;
			if (nb == ne) continue;                   // unnamed: not kept as a child (sXML.hpp:120-121)
		}
		const char *kb, *ke, *vb, *ve;
		const char *p3b = nullptr, *p3e = nullptr, *db = nullptr, *de = nullptr, *tb = nullptr, *te = nullptr;
		// a child named "x/" (self-closed, no space, no properties) returns early from sXML::process but IS kept as a
		// child, i.e. it is a point with empty p3d/desc/desc_type
		while (opened && next_property(c, kb, ke, vb, ve)) {
			if (key_is(kb, ke, "p3d")) { p3b = vb; p3e = ve; }
			else if (key_is(kb, ke, "desc")) { db = vb; de = ve; }
			else if (key_is(kb, ke, "desc_type")) { tb = vb; te = ve; }
		}
		// the point
		if (tb) type_key = unescape(tb, te); else type_key.clear();
		if (!have_cur || type_key != cur_key) {
			cur = &m.ips[type_key]; cur_key = type_key; have_cur = true;
			if (cur->desc.capacity() == 0) {                 // first point of this type: room for what the rest of the file can hold
				const size_t rest = (size_t)(c.end - c.p);
				cur->desc.reserve(rest / 4 + 16);
				cur->xyz.reserve(rest / 128 + 16);
			}
		}
		p3.clear();
		if (p3b) read_floats(p3b, p3e, p3, 3);
		while (p3.size() < 3) p3.push_back(0.f);      // the reference leaves missing coordinates uninitialised
		for (int i = 0; i < 3; i++) {
			cur->xyz.push_back(p3[i]);
			m.bbox[i] = std::min(m.bbox[i], p3[i]);
			m.bbox[3 + i] = std::max(m.bbox[3 + i], p3[i]);
		}
		if (db) read_floats(db, de, cur->desc, (size_t)-1);
		cur->desc_off.push_back((uint32_t)cur->desc.size());
		if (!opened) continue;
		skip_children(c);                             // Observation entries of a full export
		finish_element(c);
	}
	finish_element(c);
}

void parse_model(const char *data, size_t len, ModelRec &m) {
	Cursor c{data, data + len};
	const char *nb, *ne;
	if (!open_tag(c, nb, ne)) throw ParseError{"no root element"};
	const char *kb, *ke, *vb, *ve;
	while (next_property(c, kb, ke, vb, ve))
		if (key_is(kb, ke, "name")) m.name = unescape(vb, ve);
	while (!c.eof() && *c.p == '>') {
		if (!open_tag(c, nb, ne)) {
			if (is_closing(nb, ne)) break;
			continue;
		}
		const bool is_points = key_is(nb, ne, "Points");
		while (next_property(c, kb, ke, vb, ve)) {}
		if (is_points) {
			m.ips.clear();                            // the LAST <Points> child wins (moped.cpp:110-113)
			for (int i = 0; i < 3; i++) { m.bbox[i] = 10E10f; m.bbox[3 + i] = -10E10f; }
			m.has_points = true;
			parse_points(c, m);
		} else {
			skip_children(c);
			finish_element(c);
			continue;
		}
	}
}

struct MappedFile {
	const char *data = nullptr;
	size_t len = 0;
	int fd = -1;
	bool mapped = false;
	std::string owned;
	bool open(const char *path, std::string &err) {
		fd = ::open(path, O_RDONLY);
		if (fd < 0) { err = std::string("cannot open ") + path + ": " + strerror(errno); return false; }
		struct stat st;
		if (fstat(fd, &st) != 0) { err = std::string("cannot stat ") + path; ::close(fd); fd = -1; return false; }
		len = (size_t)st.st_size;
		if (len == 0) { data = ""; return true; }
		void *p = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
		if (p == MAP_FAILED) { err = std::string("cannot mmap ") + path; ::close(fd); fd = -1; return false; }
		madvise(p, len, MADV_SEQUENTIAL);
		data = (const char *)p;
		mapped = true;
		return true;
	}
	~MappedFile() {
		if (mapped) munmap((void *)data, len);
		if (fd >= 0) ::close(fd);
	}
};

} // namespace

struct mc_model_db {
	std::vector<ModelRec> models;
	std::vector<Packed> packed;          // results of mc_model_db_pack stay alive until the next change
	std::string err;
};

namespace {

// MopedPimpl::addModel(SP_Model&), moped.cpp:138-149: every model with the same name is replaced in place; the model is
// ALSO appended unless the LAST model of the list has that name — the reference's loop overwrites `found` on every
// iteration (`if( (found = (m->name == model->name)) ) m = model;`), so it only remembers the last comparison. The
// model list decides the matcher's row numbering, hence the quirk is kept.
void add_model(mc_model_db *db, ModelRec &m) {
	db->packed.clear();
	bool found = false;
	for (size_t i = 0; i < db->models.size(); i++) {
		found = db->models[i].name == m.name;
		if (found) db->models[i] = m;
	}
	if (!found) db->models.push_back(m);
}

mc_status parse_buffer(mc_model_db *db, const char *data, size_t len, ModelRec &m, const char *what) {
	try {
		parse_model(data, len, m);
	} catch (const ParseError &e) {
		db->err = std::string(what) + ": " + e.what;
		return MC_ERR_ARG;
	}
	if (!m.has_points) { db->err = std::string(what) + ": no <Points> element"; return MC_ERR_ARG; }   // addModel returns "" (:115)
	return MC_OK;
}

const uint64_t kCacheMagic = 0x31424445504f4dULL;   // "MOPEDB1"

} // namespace

extern "C" {

mc_status mc_model_db_create(mc_model_db **db) {
	if (!db) return MC_ERR_ARG;
	*db = new mc_model_db;
	return MC_OK;
}

void mc_model_db_destroy(mc_model_db *db) { delete db; }

const char *mc_model_db_last_error(const mc_model_db *db) { return db ? db->err.c_str() : "null model database"; }

mc_status mc_model_db_add_xml_buffer(mc_model_db *db, const char *data, int64_t len) {
	if (!db || !data || len < 0) { if (db) db->err = "mc_model_db_add_xml_buffer: bad argument"; return MC_ERR_ARG; }
	ModelRec m;
	mc_status st = parse_buffer(db, data, (size_t)len, m, "xml buffer");
	if (st != MC_OK) return st;
	add_model(db, m);
	return MC_OK;
}

mc_status mc_model_db_add_xml_files(mc_model_db *db, const char *const *paths, int n_files, int n_threads) {
	if (!db || !paths || n_files < 0) { if (db) db->err = "mc_model_db_add_xml_files: bad argument"; return MC_ERR_ARG; }
	if (n_files == 0) return MC_OK;
	std::vector<ModelRec> recs((size_t)n_files);
	std::vector<std::string> errs((size_t)n_files);
	std::atomic<int> next(0);
	auto work = [&]() {
		for (;;) {
			const int i = next.fetch_add(1);
			if (i >= n_files) return;
			MappedFile f;
			if (!f.open(paths[i], errs[i])) continue;
			try {
				parse_model(f.data, f.len, recs[i]);
				if (!recs[i].has_points) errs[i] = std::string(paths[i]) + ": no <Points> element";
			} catch (const ParseError &e) {
				errs[i] = std::string(paths[i]) + ": " + e.what;
			}
		}
	};
	int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
	if (nt < 1) nt = 1;
	if (nt > n_files) nt = n_files;
	std::vector<std::thread> pool;
	for (int t = 1; t < nt; t++) pool.emplace_back(work);
	work();
	for (auto &t : pool) t.join();
	for (int i = 0; i < n_files; i++)
		if (!errs[i].empty()) { db->err = errs[i]; return MC_ERR_ARG; }     // nothing is added when any file fails
	for (int i = 0; i < n_files; i++) add_model(db, recs[i]);
	return MC_OK;
}

mc_status mc_model_db_add_xml_file(mc_model_db *db, const char *path) {
	const char *p[1] = { path };
	if (!path) { if (db) db->err = "mc_model_db_add_xml_file: bad argument"; return MC_ERR_ARG; }
	return mc_model_db_add_xml_files(db, p, 1, 1);
}

mc_status mc_model_db_remove(mc_model_db *db, const char *name) {          // MopedPimpl::removeModel, moped.cpp:151-158
	if (!db || !name) return MC_ERR_ARG;
	db->packed.clear();
	for (size_t i = 0; i < db->models.size();)
		if (db->models[i].name == name) db->models.erase(db->models.begin() + (long)i); else i++;
	return MC_OK;
}

int mc_model_db_n_models(const mc_model_db *db) { return db ? (int)db->models.size() : 0; }

const char *mc_model_db_model_name(const mc_model_db *db, int i) {
	if (!db || i < 0 || i >= (int)db->models.size()) return NULL;
	return db->models[(size_t)i].name.c_str();
}

mc_status mc_model_db_model_bbox(const mc_model_db *db, int i, float *bbox6) {
	if (!db || !bbox6 || i < 0 || i >= (int)db->models.size()) return MC_ERR_ARG;
	memcpy(bbox6, db->models[(size_t)i].bbox, sizeof(float) * 6);
	return MC_OK;
}

// Rows of `desc_type` in the matcher's order (MATCH_ANN_CPU::Update :85-100). desc_size = DescriptorSize of the
// MATCH stage: the first desc_size values of every descriptor are taken (`refPts[x][i] = descriptor[i]`, :95-96); a
// shorter descriptor is an error (the reference reads past the vector). normalise != 0 applies norm() to the WHOLE
// stored descriptor first (:54-57,94), with the stage class's fp32 expression (moped_b200/stages/MATCH_CUDA.hpp).
mc_status mc_model_db_pack(mc_model_db *db, const char *desc_type, int desc_size, int normalise, int64_t *n_rows, const float **desc,
                           const float **xyz, const int32_t **model_of_row, const int32_t **n_pts_per_model) {
	if (!db || !desc_type || desc_size <= 0 || !n_rows) { if (db) db->err = "mc_model_db_pack: bad argument"; return MC_ERR_ARG; }
	for (size_t k = 0; k < db->packed.size(); k++)
		if (db->packed[k].desc_type == desc_type && db->packed[k].D == desc_size && db->packed[k].normalised == (normalise != 0)) {
			const Packed &P = db->packed[k];
			*n_rows = (int64_t)P.model_of_row.size();
			if (desc) *desc = P.desc.data();
			if (xyz) *xyz = P.xyz.data();
			if (model_of_row) *model_of_row = P.model_of_row.data();
			if (n_pts_per_model) *n_pts_per_model = P.n_pts.data();
			return MC_OK;
		}
	Packed P;
	P.desc_type = desc_type; P.D = desc_size; P.normalised = normalise != 0;
	size_t rows = 0;
	for (const ModelRec &m : db->models) {
		auto it = m.ips.find(desc_type);
		const size_t n = it == m.ips.end() ? 0 : it->second.n();
		P.n_pts.push_back((int32_t)n);
		rows += n;
	}
	P.desc.resize(rows * (size_t)desc_size);
	P.xyz.resize(rows * 3);
	P.model_of_row.resize(rows);
	size_t row = 0;
	std::vector<float> tmp;
	for (size_t mi = 0; mi < db->models.size(); mi++) {
		auto it = db->models[mi].ips.find(desc_type);
		if (it == db->models[mi].ips.end()) continue;
		const PointSet &ps = it->second;
		for (size_t f = 0; f < ps.n(); f++, row++) {
			const size_t lo = ps.desc_off[f], len = ps.desc_off[f + 1] - lo;
			if (len < (size_t)desc_size) {
				db->err = "mc_model_db_pack: model '" + db->models[mi].name + "' has a '" + desc_type + "' descriptor of " + std::to_string(len) +
				          " values, fewer than DescriptorSize " + std::to_string(desc_size);
				return MC_ERR_ARG;
			}
			const float *src = ps.desc.data() + lo;
			float *dst = P.desc.data() + row * (size_t)desc_size;
			if (normalise) {
				float ss = 0;
				for (size_t k = 0; k < len; k++) ss += src[k] * src[k];
				const float inv = 1. / sqrtf(ss);
				for (int k = 0; k < desc_size; k++) dst[k] = src[k] * inv;
			} else {
				memcpy(dst, src, sizeof(float) * (size_t)desc_size);
			}
			memcpy(P.xyz.data() + row * 3, ps.xyz.data() + f * 3, sizeof(float) * 3);
			P.model_of_row[row] = (int32_t)mi;
		}
	}
	db->packed.push_back(std::move(P));
	const Packed &R = db->packed.back();
	*n_rows = (int64_t)rows;
	if (desc) *desc = R.desc.data();
	if (xyz) *xyz = R.xyz.data();
	if (model_of_row) *model_of_row = R.model_of_row.data();
	if (n_pts_per_model) *n_pts_per_model = R.n_pts.data();
	return MC_OK;
}

// pack + normalise + mc_db_upload: what `modelsUpdated` -> MATCH::Update does, ending in the device-resident database
mc_status mc_model_db_upload(mc_model_db *db, mc_ctx *ctx, const char *desc_type, int desc_size) {
	if (!db || !ctx) return MC_ERR_ARG;
	int64_t n = 0;
	const float *desc = nullptr, *xyz = nullptr;
	const int32_t *mor = nullptr;
	mc_status st = mc_model_db_pack(db, desc_type, desc_size, 1, &n, &desc, &xyz, &mor, nullptr);
	if (st != MC_OK) return st;
	if (n < 2) { db->err = "mc_model_db_upload: fewer than two rows (the reference skips matching, MATCH_ANN_CPU.hpp:102)"; return MC_ERR_STATE; }
	st = mc_db_upload(ctx, desc, xyz, mor, n, desc_size, (int)db->models.size(), 0);
	if (st != MC_OK) db->err = std::string("mc_db_upload: ") + mc_last_error(ctx);
	return st;
}

// ---- binary cache: the parsed models (all descriptor types, unnormalised), little-endian ----
//   u64 magic, u32 n_models; per model: u32 name_len, name, 6 f32 bbox, u32 n_types;
//   per type: u32 key_len, key, u64 n_points, u64 n_desc_values, xyz[3n] f32, desc_off[n+1] u32, desc[] f32
mc_status mc_model_db_save(const mc_model_db *db_c, const char *path) {
	mc_model_db *db = const_cast<mc_model_db *>(db_c);
	if (!db || !path) return MC_ERR_ARG;
	FILE *f = fopen(path, "wb");
	if (!f) { db->err = std::string("cannot write ") + path; return MC_ERR_ARG; }
	bool ok = true;
	auto w = [&](const void *p, size_t n) { if (n && fwrite(p, 1, n, f) != n) ok = false; };
	auto w32 = [&](uint32_t v) { w(&v, 4); };
	auto w64 = [&](uint64_t v) { w(&v, 8); };
	w64(kCacheMagic);
	w32((uint32_t)db->models.size());
	for (const ModelRec &m : db->models) {
		w32((uint32_t)m.name.size()); w(m.name.data(), m.name.size());
		w(m.bbox, sizeof m.bbox);
		w32((uint32_t)m.ips.size());
		for (const auto &kv : m.ips) {
			w32((uint32_t)kv.first.size()); w(kv.first.data(), kv.first.size());
			w64(kv.second.n()); w64(kv.second.desc.size());
			w(kv.second.xyz.data(), kv.second.xyz.size() * 4);
			w(kv.second.desc_off.data(), kv.second.desc_off.size() * 4);
			w(kv.second.desc.data(), kv.second.desc.size() * 4);
		}
	}
	if (fclose(f) != 0) ok = false;
	if (!ok) { db->err = std::string("short write to ") + path; return MC_ERR_ARG; }
	return MC_OK;
}

mc_status mc_model_db_load(mc_model_db *db, const char *path) {
	if (!db || !path) return MC_ERR_ARG;
	MappedFile f;
	if (!f.open(path, db->err)) return MC_ERR_ARG;
	const char *p = f.data, *end = f.data + f.len;
	bool ok = true;
	auto need = [&](size_t n) { if ((size_t)(end - p) < n) ok = false; return ok; };
	auto r32 = [&]() { uint32_t v = 0; if (need(4)) { memcpy(&v, p, 4); p += 4; } return v; };
	auto r64 = [&]() { uint64_t v = 0; if (need(8)) { memcpy(&v, p, 8); p += 8; } return v; };
	if (r64() != kCacheMagic || !ok) { db->err = std::string(path) + ": not a moped model cache"; return MC_ERR_ARG; }
	const uint32_t nm = r32();
	std::vector<ModelRec> recs;
	for (uint32_t i = 0; ok && i < nm; i++) {
		ModelRec m;
		const uint32_t nl = r32();
		if (!need(nl)) break;
		m.name.assign(p, nl); p += nl;
		if (!need(sizeof m.bbox)) break;
		memcpy(m.bbox, p, sizeof m.bbox); p += sizeof m.bbox;
		const uint32_t nt = r32();
		for (uint32_t t = 0; ok && t < nt; t++) {
			const uint32_t kl = r32();
			if (!need(kl)) break;
			std::string key(p, kl); p += kl;
			const uint64_t np = r64(), nd = r64();
			if (!ok || np > (1ull << 40) || nd > (1ull << 40)) { ok = false; break; }
			if (!need(np * 12 + (np + 1) * 4 + nd * 4)) break;
			PointSet &ps = m.ips[key];
			ps.xyz.resize(np * 3); memcpy(ps.xyz.data(), p, np * 12); p += np * 12;
			ps.desc_off.resize(np + 1); memcpy(ps.desc_off.data(), p, (np + 1) * 4); p += (np + 1) * 4;
			ps.desc.resize(nd); memcpy(ps.desc.data(), p, nd * 4); p += nd * 4;
			if (ps.desc_off[0] != 0 || ps.desc_off[np] != nd) ok = false;
			for (uint64_t k = 0; ok && k < np; k++)                  // offsets of an untrusted file: non-decreasing and inside the descriptor array
				if (ps.desc_off[k + 1] < ps.desc_off[k] || ps.desc_off[k + 1] > nd) ok = false;
		}
		m.has_points = true;
		recs.push_back(std::move(m));
	}
	if (!ok || recs.size() != nm) { db->err = std::string(path) + ": truncated or corrupt model cache"; return MC_ERR_ARG; }
	for (ModelRec &m : recs) add_model(db, m);
	return MC_OK;
}

} // extern "C"
