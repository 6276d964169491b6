// OpenFOAM-free case I/O (see foam_io.h).  Formats handled are the ones the reference's examples use
// (SURVEY.md section 8(c)): ASCII polyMesh, ASCII dictionaries, uniform / nonuniform internalField.
#include "foam_io.h"

#include <sys/stat.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace lfm {

namespace {

[[noreturn]] void fail(const std::string& msg) { throw std::runtime_error("lfm::foam_io: " + msg); }

std::string slurp(const std::string& path) {
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) fail("cannot open " + path);
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	std::string s;
	s.resize((size_t)n);
	if (n > 0 && fread(&s[0], 1, (size_t)n, f) != (size_t)n) {
		fclose(f);
		fail("short read on " + path);
	}
	fclose(f);
	return s;
}

bool fileExists(const std::string& path) {
	struct stat st;
	return stat(path.c_str(), &st) == 0;
}

// Replaces // and /* */ comments by blanks (strings are respected).
void stripComments(std::string& s) {
	const size_t n = s.size();
	size_t i = 0;
	while (i < n) {
		char c = s[i];
		if (c == '"') {
			i++;
			while (i < n && s[i] != '"') {
				if (s[i] == '\\') i++;
				i++;
			}
			i++;
		} else if (c == '/' && i + 1 < n && s[i + 1] == '/') {
			while (i < n && s[i] != '\n') s[i++] = ' ';
		} else if (c == '/' && i + 1 < n && s[i + 1] == '*') {
			s[i++] = ' ';
			s[i++] = ' ';
			while (i < n && !(s[i] == '*' && i + 1 < n && s[i + 1] == '/')) {
				if (s[i] != '\n') s[i] = ' ';
				i++;
			}
			if (i < n) {
				s[i++] = ' ';
				if (i < n) s[i++] = ' ';
			}
		} else {
			i++;
		}
	}
}

struct Tokenizer {
	const char* p;
	const char* end;
	explicit Tokenizer(const std::string& s) : p(s.data()), end(s.data() + s.size()) {}
	static bool isPunct(char c) { return c == '{' || c == '}' || c == '(' || c == ')' || c == ';' || c == '[' || c == ']'; }
	void skipWs() {
		while (p < end && (unsigned char)*p <= ' ') p++;
	}
	// returns false at EOF
	bool next(std::string& tok) {
		skipWs();
		if (p >= end) return false;
		char c = *p;
		if (c == '"') {
			const char* s = p++;
			while (p < end && *p != '"') {
				if (*p == '\\') p++;
				p++;
			}
			if (p < end) p++;
			tok.assign(s, p);
			return true;
		}
		if (isPunct(c)) {
			tok.assign(1, c);
			p++;
			return true;
		}
		if (c == '#') {  // directive: swallow the line
			const char* s = p;
			while (p < end && *p != '\n') p++;
			tok.assign(s, p);
			return true;
		}
		// word / number; attached balanced (...) stays inside the word: div(tauMC), reconstruct(p)
		const char* s = p;
		int depth = 0;
		while (p < end) {
			char d = *p;
			if ((unsigned char)d <= ' ') break;
			if (d == '(') {
				depth++;
			} else if (d == ')') {
				if (depth == 0) break;
				depth--;
			} else if (depth == 0 && (d == '{' || d == '}' || d == ';' || d == '[' || d == ']' || d == '"')) {
				break;
			}
			p++;
		}
		tok.assign(s, p);
		return true;
	}
};

void parseDictBody(Tokenizer& tk, Dict& d, bool top) {
	std::string tok;
	while (tk.next(tok)) {
		if (tok == "}") {
			if (top) fail("unbalanced '}' in dictionary");
			return;
		}
		if (tok == ";") continue;
		if (tok[0] == '#') continue;
		DictEntry e;
		e.key = tok;
		if (e.key.size() >= 2 && e.key.front() == '"' && e.key.back() == '"') e.key = e.key.substr(1, e.key.size() - 2);
		std::string t2;
		if (!tk.next(t2)) {
			d.entries.push_back(e);
			return;
		}
		if (t2 == "{") {
			e.is_dict = true;
			e.sub = std::make_shared<Dict>();
			parseDictBody(tk, *e.sub, false);
			d.entries.push_back(std::move(e));
			continue;
		}
		int paren = 0, brace = 0;
		std::string t = t2;
		while (true) {
			if (t == ";" && paren == 0 && brace == 0) break;
			if (t == "(" || t == "[") paren++;
			if (t == ")" || t == "]") paren--;
			if (t == "{") brace++;
			if (t == "}") {
				if (brace == 0) fail("unexpected '}' in value of key " + e.key);
				brace--;
			}
			e.tokens.push_back(t);
			if (!tk.next(t)) break;
		}
		d.entries.push_back(std::move(e));
	}
	if (!top) fail("missing '}' at end of dictionary");
}

}  // namespace

// ---------------------------------------------------------------------------------------------
const DictEntry* Dict::find(const std::string& key) const {
	// last definition wins (OpenFOAM merges duplicates; later plain entries overwrite)
	for (size_t i = entries.size(); i-- > 0;)
		if (entries[i].key == key) return &entries[i];
	return nullptr;
}

const Dict& Dict::optionalSubDict(const std::string& key) const {
	const DictEntry* e = find(key);
	if (e && e->is_dict) return *e->sub;
	return *this;
}

bool parseSwitch(const std::string& t, bool& out) {
	if (t == "true" || t == "yes" || t == "on" || t == "y" || t == "t" || t == "any") {
		out = true;
		return true;
	}
	if (t == "false" || t == "no" || t == "off" || t == "n" || t == "f" || t == "none") {
		out = false;
		return true;
	}
	return false;
}

double Dict::scalar(const std::string& key) const {
	const DictEntry* e = find(key);
	if (!e || e->is_dict || e->tokens.empty()) fail("keyword " + key + " is undefined in dictionary");
	char* endp = nullptr;
	double v = strtod(e->tokens[0].c_str(), &endp);
	if (endp == e->tokens[0].c_str()) {
		bool b;
		if (parseSwitch(e->tokens[0], b)) return b ? 1.0 : 0.0;
		fail("keyword " + key + " is not a scalar: " + e->tokens[0]);
	}
	return v;
}
double Dict::scalarOr(const std::string& key, double def) const { return has(key) ? scalar(key) : def; }

bool Dict::boolean(const std::string& key) const {
	const DictEntry* e = find(key);
	if (!e || e->is_dict || e->tokens.empty()) fail("keyword " + key + " is undefined in dictionary");
	bool b;
	if (parseSwitch(e->tokens[0], b)) return b;
	char* endp = nullptr;
	double v = strtod(e->tokens[0].c_str(), &endp);
	if (endp == e->tokens[0].c_str()) fail("keyword " + key + " is not a bool: " + e->tokens[0]);
	return v != 0.0;
}
bool Dict::booleanOr(const std::string& key, bool def) const { return has(key) ? boolean(key) : def; }

std::string Dict::word(const std::string& key) const {
	const DictEntry* e = find(key);
	if (!e || e->is_dict || e->tokens.empty()) fail("keyword " + key + " is undefined in dictionary");
	std::string w = e->tokens[0];
	if (w.size() >= 2 && w.front() == '"' && w.back() == '"') w = w.substr(1, w.size() - 2);
	return w;
}
std::string Dict::wordOr(const std::string& key, const std::string& def) const { return has(key) ? word(key) : def; }

std::string Dict::valueString(const std::string& key) const {
	const DictEntry* e = find(key);
	if (!e || e->is_dict) fail("keyword " + key + " is undefined in dictionary");
	std::string s;
	for (size_t i = 0; i < e->tokens.size(); i++) {
		if (i) s += ' ';
		s += e->tokens[i];
	}
	return s;
}

Dict parseDictString(const std::string& text) {
	std::string s = text;
	stripComments(s);
	Tokenizer tk(s);
	Dict d;
	parseDictBody(tk, d, true);
	return d;
}

Dict parseDictFile(const std::string& path) { return parseDictString(slurp(path)); }

// ---------------------------------------------------------------------------------------------
// polyMesh reading
// ---------------------------------------------------------------------------------------------
namespace {

// What the FoamFile{...} header says about the encoding of the lists that follow.  OpenFOAM's binary stream format
// (IOstreamOption::BINARY; the 3D examples say `writeFormat binary`, and decomposePar writes what the case says) keeps
// every token ASCII except the contents of contiguous lists: `N(` + N raw elements + `)`, little endian, element sizes
// given by `arch "LSB;label=32;scalar=64"`.
struct FoamHeader {
	bool binary = false;
	int labelBytes = 4, scalarBytes = 8;
	size_t bodyStart = 0;      // offset behind the header's closing brace in the RAW (comments not stripped) text
};

FoamHeader parseHeader(const std::string& raw, const std::string& path) {
	FoamHeader h;
	size_t pos = raw.find("FoamFile");
	if (pos == std::string::npos) return h;
	size_t open = raw.find('{', pos);
	size_t close = open == std::string::npos ? std::string::npos : raw.find('}', open);
	if (open == std::string::npos || close == std::string::npos) fail("bad FoamFile header in " + path);
	h.bodyStart = close + 1;
	const std::string hdr = raw.substr(open + 1, close - open - 1);
	size_t f = hdr.find("format");
	if (f != std::string::npos) {
		size_t semi = hdr.find(';', f), b = hdr.find("binary", f);
		h.binary = b != std::string::npos && b < semi;
	}
	size_t a = hdr.find("arch");
	if (a != std::string::npos) {
		size_t semi = hdr.find("\";", a);
		const std::string arch = hdr.substr(a, semi == std::string::npos ? std::string::npos : semi - a);
		if (arch.find("MSB") != std::string::npos) fail("big-endian binary files are not supported: " + path);
		size_t l = arch.find("label="), sc = arch.find("scalar=");
		if (l != std::string::npos) h.labelBytes = atoi(arch.c_str() + l + 6) / 8;
		if (sc != std::string::npos) h.scalarBytes = atoi(arch.c_str() + sc + 7) / 8;
	}
	if (h.binary && ((h.labelBytes != 4 && h.labelBytes != 8) || (h.scalarBytes != 4 && h.scalarBytes != 8))) fail("unsupported arch in " + path);
	return h;
}

// Positions `p` after the FoamFile{...} header (if any) of comment-stripped text.
const char* skipHeader(const std::string& s, const std::string& path) {
	size_t pos = s.find("FoamFile");
	if (pos == std::string::npos) return s.data();
	size_t open = s.find('{', pos);
	size_t close = s.find('}', open);
	if (open == std::string::npos || close == std::string::npos) fail("bad FoamFile header in " + path);
	return s.data() + close + 1;
}

inline void skipWs(const char*& p, const char* end) {
	while (p < end && (unsigned char)*p <= ' ') p++;
}

long readCount(const char*& p, const char* end, const std::string& path) {
	skipWs(p, end);
	char* e = nullptr;
	long n = strtol(p, &e, 10);
	if (e == p) fail("expected a list size in " + path);
	p = e;
	skipWs(p, end);
	if (p >= end || *p != '(') fail("expected '(' after list size in " + path);
	p++;
	return n;
}

// ---- binary lists: the text between lists is not comment-stripped (the raw bytes may hold any pattern) ----
inline void skipWsComments(const char*& p, const char* end) {
	for (;;) {
		while (p < end && (unsigned char)*p <= ' ') p++;
		if (p + 1 < end && p[0] == '/' && p[1] == '/') {
			while (p < end && *p != '\n') p++;
		} else if (p + 1 < end && p[0] == '/' && p[1] == '*') {
			p += 2;
			while (p + 1 < end && !(p[0] == '*' && p[1] == '/')) p++;
			p = p + 2 <= end ? p + 2 : end;
		} else {
			return;
		}
	}
}
// `N(` : returns N with p on the first raw byte
long readCountRaw(const char*& p, const char* end, const std::string& path) {
	skipWsComments(p, end);
	char* e = nullptr;
	long n = strtol(p, &e, 10);
	if (e == p || n < 0) fail("expected a list size in " + path);
	p = e;
	skipWsComments(p, end);
	if (p >= end || *p != '(') fail("expected '(' after list size in " + path);
	p++;
	return n;
}
inline void closeRaw(const char*& p, const char* end, const std::string& path) {
	if (p >= end || *p != ')') fail("expected ')' behind the binary block in " + path);
	p++;
}
std::vector<int> readRawLabels(const char*& p, const char* end, int labelBytes, const std::string& path) {
	const long n = readCountRaw(p, end, path);
	if ((size_t)(end - p) < (size_t)n * labelBytes) fail("truncated binary label list in " + path);
	std::vector<int> v((size_t)n);
	if (labelBytes == 4) {
		if (n) memcpy(v.data(), p, (size_t)n * 4);
	} else {
		for (long i = 0; i < n; i++) {
			int64_t x;
			memcpy(&x, p + (size_t)i * 8, 8);
			v[(size_t)i] = (int)x;
		}
	}
	p += (size_t)n * labelBytes;
	closeRaw(p, end, path);
	return v;
}
// n_expected < 0: any size; returns the scalars widened to double
std::vector<double> readRawScalars(const char*& p, const char* end, int scalarBytes, int nComp, long n_expected, const std::string& path) {
	const long n = readCountRaw(p, end, path);
	if (n_expected >= 0 && n != n_expected) fail("list size mismatch in " + path);
	const size_t cnt = (size_t)n * nComp;
	if ((size_t)(end - p) < cnt * scalarBytes) fail("truncated binary scalar list in " + path);
	std::vector<double> v(cnt);
	if (scalarBytes == 8) {
		if (cnt) memcpy(v.data(), p, cnt * 8);
	} else {
		for (size_t i = 0; i < cnt; i++) {
			float x;
			memcpy(&x, p + i * 4, 4);
			v[i] = (double)x;
		}
	}
	p += cnt * scalarBytes;
	closeRaw(p, end, path);
	return v;
}

std::vector<int> readLabelList(const std::string& path) {
	std::string s = slurp(path);
	const FoamHeader hd = parseHeader(s, path);
	if (hd.binary) {
		const char* p = s.data() + hd.bodyStart;
		return readRawLabels(p, s.data() + s.size(), hd.labelBytes, path);
	}
	stripComments(s);
	const char* end = s.data() + s.size();
	const char* p = skipHeader(s, path);
	long n = readCount(p, end, path);
	std::vector<int> v((size_t)n);
	for (long i = 0; i < n; i++) {
		char* e = nullptr;
		long x = strtol(p, &e, 10);
		if (e == p) fail("bad label in " + path);
		v[(size_t)i] = (int)x;
		p = e;
	}
	return v;
}

void writeHeader(FILE* f, const char* cls, const char* object, const char* location, const char* note = nullptr) {
	fprintf(f, "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       %s;\n", cls);
	if (note) fprintf(f, "    note        \"%s\";\n", note);
	if (location) fprintf(f, "    location    \"%s\";\n", location);
	fprintf(f, "    object      %s;\n}\n\n", object);
}

void writeLabelList(const std::string& path, const char* object, const std::vector<int>& v, const char* note = nullptr) {
	FILE* f = fopen(path.c_str(), "w");
	if (!f) fail("cannot write " + path);
	writeHeader(f, "labelList", object, "constant/polyMesh", note);
	fprintf(f, "%zu\n(\n", v.size());
	for (int x : v) fprintf(f, "%d\n", x);
	fprintf(f, ")\n");
	fclose(f);
}

}  // namespace

int PolyMesh::nNonProcessor() const {
	int n = 0;
	for (const Patch& p : patches) {
		if (p.isProcessor()) break;
		n++;
	}
	return n;
}

int PolyMesh::whichPatch(int face) const {
	if (face < nInternalFaces()) return -1;
	for (size_t i = 0; i < patches.size(); i++)
		if (face >= patches[i].startFace && face < patches[i].startFace + patches[i].nFaces) return (int)i;
	return -1;
}

void PolyMesh::resolvePatches() {
	for (Patch& p : patches) {
		p.neighbPatchID = -1;
		p.referPatchID = -1;
		if (p.isCyclic())
			for (size_t j = 0; j < patches.size(); j++)
				if (patches[j].name == p.neighbourPatch) p.neighbPatchID = (int)j;
		if (p.isProcessorCyclic()) {
			for (size_t j = 0; j < patches.size(); j++)
				if (patches[j].name == p.referPatch) p.referPatchID = (int)j;
			if (p.referPatchID < 0) fail("processorCyclic patch " + p.name + " refers to the unknown patch " + p.referPatch);
		}
	}
}

// The signed tag the reference matches MPI faces with (polyMeshReaderOF.cpp:460-475 getBoundaryTag = tag() * (2 owner() - 1)).
// Plain processor patches: tag() == UPstream::msgType() == 1.  processorCyclic patches: OpenFOAM hashes the name of the
// referred cyclic patch on the owner side and of that patch's neighbour on the other side, so that both sides get the same
// value; only that agreement (and a value outside [-1, 1], which switches getFaceId to the index inside the patch) matters
// to the reference, so the hash here is FNV-1a folded into [2, 32767] rather than OpenFOAM's Jenkins hash.
int PolyMesh::processorTag(int b) const {
	const Patch& p = patches[(size_t)b];
	if (!p.isProcessor()) return 0;
	const int sign = p.myProcNo < p.neighbProcNo ? 1 : -1;   // processorPolyPatch::owner()
	if (!p.isProcessorCyclic()) return sign;
	if (p.tag > 1) return sign * p.tag;
	const Patch& ref = patches[(size_t)p.referPatchID];
	if (sign < 0 && ref.neighbPatchID < 0) fail("cyclic patch " + ref.name + " has no neighbourPatch");
	const std::string& name = sign > 0 ? ref.name : patches[(size_t)ref.neighbPatchID].name;
	uint32_t h = 2166136261u;
	for (unsigned char ch : name) h = (h ^ ch) * 16777619u;
	return sign * (int)(2u + h % 32766u);
}

void PolyMesh::buildCells() {
	// OpenFOAM primitiveMesh::calcCells: owner pass over all faces, then neighbour pass
	int nc = 0;
	for (int o : owner) nc = o + 1 > nc ? o + 1 : nc;
	for (int n : neighbour) nc = n + 1 > nc ? n + 1 : nc;
	if (nCells < nc) nCells = nc;
	cellFaceOffsets.assign((size_t)nCells + 1, 0);
	for (int o : owner) cellFaceOffsets[(size_t)o + 1]++;
	for (int n : neighbour) cellFaceOffsets[(size_t)n + 1]++;
	for (int c = 0; c < nCells; c++) cellFaceOffsets[(size_t)c + 1] += cellFaceOffsets[(size_t)c];
	cellFaces.assign((size_t)cellFaceOffsets[(size_t)nCells], -1);
	std::vector<int> fill(cellFaceOffsets.begin(), cellFaceOffsets.end() - 1);
	for (int f = 0; f < nFaces(); f++) cellFaces[(size_t)fill[(size_t)owner[(size_t)f]]++] = f;
	for (int f = 0; f < nInternalFaces(); f++) cellFaces[(size_t)fill[(size_t)neighbour[(size_t)f]]++] = f;
}

void PolyMesh::computeGeometry() {
	const int nf = nFaces();
	faceAreas.assign((size_t)nf * 3, 0.0);
	faceCentres.assign((size_t)nf * 3, 0.0);
	const double ROOTVSMALL = 1.0e-150;
	const double VSMALL = 1.0e-300;
	// primitiveMeshFaceCentresAndAreas (OpenFOAM, classic triangle-fan form)
	for (int f = 0; f < nf; f++) {
		const int* fp = &facePoints[(size_t)faceOffsets[(size_t)f]];
		const int np = facePointCount(f);
		double* A = &faceAreas[(size_t)f * 3];
		double* C = &faceCentres[(size_t)f * 3];
		if (np == 3) {
			const double* a = &points[(size_t)fp[0] * 3];
			const double* b = &points[(size_t)fp[1] * 3];
			const double* c = &points[(size_t)fp[2] * 3];
			for (int k = 0; k < 3; k++) C[k] = (1.0 / 3.0) * (a[k] + b[k] + c[k]);
			double u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
			double v[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
			A[0] = 0.5 * (u[1] * v[2] - u[2] * v[1]);
			A[1] = 0.5 * (u[2] * v[0] - u[0] * v[2]);
			A[2] = 0.5 * (u[0] * v[1] - u[1] * v[0]);
		} else {
			double fc[3] = {points[(size_t)fp[0] * 3], points[(size_t)fp[0] * 3 + 1], points[(size_t)fp[0] * 3 + 2]};
			for (int i = 1; i < np; i++)
				for (int k = 0; k < 3; k++) fc[k] += points[(size_t)fp[i] * 3 + k];
			for (int k = 0; k < 3; k++) fc[k] /= np;
			double sumN[3] = {0, 0, 0}, sumAc[3] = {0, 0, 0}, sumA = 0.0;
			for (int i = 0; i < np; i++) {
				const double* pi = &points[(size_t)fp[i] * 3];
				const double* pn = &points[(size_t)fp[(i + 1) % np] * 3];
				double c[3], u[3], v[3], n[3];
				for (int k = 0; k < 3; k++) {
					c[k] = pi[k] + pn[k] + fc[k];
					u[k] = pn[k] - pi[k];
					v[k] = fc[k] - pi[k];
				}
				n[0] = u[1] * v[2] - u[2] * v[1];
				n[1] = u[2] * v[0] - u[0] * v[2];
				n[2] = u[0] * v[1] - u[1] * v[0];
				double a = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
				for (int k = 0; k < 3; k++) {
					sumN[k] += n[k];
					sumAc[k] += a * c[k];
				}
				sumA += a;
			}
			if (sumA < ROOTVSMALL) {
				for (int k = 0; k < 3; k++) {
					C[k] = fc[k];
					A[k] = 0.0;
				}
			} else {
				for (int k = 0; k < 3; k++) {
					C[k] = (1.0 / 3.0) * sumAc[k] / sumA;
					A[k] = 0.5 * sumN[k];
				}
			}
		}
	}
	// primitiveMeshCellCentresAndVols
	cellCentres.assign((size_t)nCells * 3, 0.0);
	cellVolumes.assign((size_t)nCells, 0.0);
	std::vector<double> cEst((size_t)nCells * 3, 0.0);
	std::vector<int> nCellFaces((size_t)nCells, 0);
	for (int f = 0; f < nf; f++) {
		int o = owner[(size_t)f];
		for (int k = 0; k < 3; k++) cEst[(size_t)o * 3 + k] += faceCentres[(size_t)f * 3 + k];
		nCellFaces[(size_t)o]++;
	}
	for (int f = 0; f < nInternalFaces(); f++) {
		int n = neighbour[(size_t)f];
		for (int k = 0; k < 3; k++) cEst[(size_t)n * 3 + k] += faceCentres[(size_t)f * 3 + k];
		nCellFaces[(size_t)n]++;
	}
	for (int c = 0; c < nCells; c++)
		for (int k = 0; k < 3; k++) cEst[(size_t)c * 3 + k] /= nCellFaces[(size_t)c];
	for (int f = 0; f < nf; f++) {
		int o = owner[(size_t)f];
		const double* A = &faceAreas[(size_t)f * 3];
		const double* C = &faceCentres[(size_t)f * 3];
		const double* E = &cEst[(size_t)o * 3];
		double pyr3Vol = A[0] * (C[0] - E[0]) + A[1] * (C[1] - E[1]) + A[2] * (C[2] - E[2]);
		for (int k = 0; k < 3; k++) cellCentres[(size_t)o * 3 + k] += pyr3Vol * ((3.0 / 4.0) * C[k] + (1.0 / 4.0) * E[k]);
		cellVolumes[(size_t)o] += pyr3Vol;
	}
	for (int f = 0; f < nInternalFaces(); f++) {
		int n = neighbour[(size_t)f];
		const double* A = &faceAreas[(size_t)f * 3];
		const double* C = &faceCentres[(size_t)f * 3];
		const double* E = &cEst[(size_t)n * 3];
		double pyr3Vol = A[0] * (E[0] - C[0]) + A[1] * (E[1] - C[1]) + A[2] * (E[2] - C[2]);
		for (int k = 0; k < 3; k++) cellCentres[(size_t)n * 3 + k] += pyr3Vol * ((3.0 / 4.0) * C[k] + (1.0 / 4.0) * E[k]);
		cellVolumes[(size_t)n] += pyr3Vol;
	}
	for (int c = 0; c < nCells; c++) {
		if (std::fabs(cellVolumes[(size_t)c]) > VSMALL)
			for (int k = 0; k < 3; k++) cellCentres[(size_t)c * 3 + k] /= cellVolumes[(size_t)c];
		else
			for (int k = 0; k < 3; k++) cellCentres[(size_t)c * 3 + k] = cEst[(size_t)c * 3 + k];
		cellVolumes[(size_t)c] *= (1.0 / 3.0);
	}
}

PolyMesh readPolyMesh(const std::string& dir) {
	PolyMesh m;
	{  // points
		std::string path = dir + "/points";
		std::string s = slurp(path);
		const FoamHeader hd = parseHeader(s, path);
		if (hd.binary) {
			const char* pb = s.data() + hd.bodyStart;
			m.points = readRawScalars(pb, s.data() + s.size(), hd.scalarBytes, 3, -1, path);
			s.clear();
		} else
			stripComments(s);
		const char* end = s.data() + s.size();
		const char* p = hd.binary ? end : skipHeader(s, path);
		long n = hd.binary ? 0 : readCount(p, end, path);
		if (!hd.binary) m.points.resize((size_t)n * 3);
		for (long i = 0; i < n; i++) {
			skipWs(p, end);
			if (p >= end || *p != '(') fail("expected '(' in " + path);
			p++;
			for (int k = 0; k < 3; k++) {
				char* e = nullptr;
				m.points[(size_t)i * 3 + k] = strtod(p, &e);
				if (e == p) fail("bad coordinate in " + path);
				p = e;
			}
			skipWs(p, end);
			if (p >= end || *p != ')') fail("expected ')' in " + path);
			p++;
		}
	}
	{  // faces
		std::string path = dir + "/faces";
		std::string s = slurp(path);
		const FoamHeader hd = parseHeader(s, path);
		if (hd.binary) {
			// class faceCompactList: the offsets (nFaces + 1 labels) followed by the concatenated point labels
			if (s.substr(0, hd.bodyStart).find("faceCompactList") == std::string::npos) fail("binary faces file is not a faceCompactList: " + path);
			const char* pb = s.data() + hd.bodyStart;
			m.faceOffsets = readRawLabels(pb, s.data() + s.size(), hd.labelBytes, path);
			m.facePoints = readRawLabels(pb, s.data() + s.size(), hd.labelBytes, path);
			if (m.faceOffsets.empty() || m.faceOffsets.front() != 0 || (size_t)m.faceOffsets.back() != m.facePoints.size()) fail("inconsistent faceCompactList in " + path);
			s.clear();
		} else {
			stripComments(s);
			if (s.find("faceCompactList") != std::string::npos) fail("ascii faceCompactList is not supported: " + path);
		}
		const char* end = s.data() + s.size();
		const char* p = hd.binary ? end : skipHeader(s, path);
		long n = hd.binary ? 0 : readCount(p, end, path);
		if (!hd.binary) {
			m.faceOffsets.resize((size_t)n + 1);
			m.facePoints.reserve((size_t)n * 4);
			m.faceOffsets[0] = 0;
		}
		for (long i = 0; i < n; i++) {
			char* e = nullptr;
			long np = strtol(p, &e, 10);
			if (e == p) fail("bad face size in " + path);
			p = e;
			skipWs(p, end);
			if (p >= end || *p != '(') fail("expected '(' in " + path);
			p++;
			for (long k = 0; k < np; k++) {
				long v = strtol(p, &e, 10);
				if (e == p) fail("bad face label in " + path);
				m.facePoints.push_back((int)v);
				p = e;
			}
			skipWs(p, end);
			if (p >= end || *p != ')') fail("expected ')' in " + path);
			p++;
			m.faceOffsets[(size_t)i + 1] = (int)m.facePoints.size();
		}
	}
	m.owner = readLabelList(dir + "/owner");
	m.neighbour = readLabelList(dir + "/neighbour");
	if (m.owner.size() + 1 != m.faceOffsets.size()) fail("owner/faces size mismatch in " + dir);
	{  // boundary
		std::string path = dir + "/boundary";
		std::string s = slurp(path);
		stripComments(s);
		const char* end = s.data() + s.size();
		const char* p = skipHeader(s, path);
		long n = readCount(p, end, path);
		std::string body(p, end);
		size_t last = body.rfind(')');
		if (last == std::string::npos) fail("missing ')' in " + path);
		body.resize(last);
		Dict d = parseDictString(body);
		if ((long)d.entries.size() != n) fail("boundary patch count mismatch in " + path);
		for (const DictEntry& e : d.entries) {
			if (!e.is_dict) fail("bad patch entry " + e.key + " in " + path);
			Patch pt;
			pt.name = e.key;
			pt.type = e.sub->word("type");
			pt.nFaces = (int)e.sub->scalar("nFaces");
			pt.startFace = (int)e.sub->scalar("startFace");
			pt.neighbourPatch = e.sub->wordOr("neighbourPatch", "");
			pt.myProcNo = (int)e.sub->scalarOr("myProcNo", -1);
			pt.neighbProcNo = (int)e.sub->scalarOr("neighbProcNo", -1);
			pt.referPatch = e.sub->wordOr("referPatch", "");
			pt.tag = (int)e.sub->scalarOr("tag", -1);
			m.patches.push_back(pt);
		}
	}
	auto optional = [&](const char* name, std::vector<int>& dst) {
		std::string path = dir + "/" + name;
		if (fileExists(path)) dst = readLabelList(path);
	};
	optional("faceProcAddressing", m.faceProcAddressing);
	optional("cellProcAddressing", m.cellProcAddressing);
	optional("pointProcAddressing", m.pointProcAddressing);
	optional("boundaryProcAddressing", m.boundaryProcAddressing);
	optional("cellSubmesh", m.cellSubmesh);
	m.finalize();
	return m;
}

void writePolyMesh(const PolyMesh& m, const std::string& dir) {
	char note[256];
	snprintf(note, sizeof note, "nPoints:%d  nCells:%d  nFaces:%d  nInternalFaces:%d", m.nPoints(), m.nCells, m.nFaces(),
	         m.nInternalFaces());
	{
		std::string path = dir + "/points";
		FILE* f = fopen(path.c_str(), "w");
		if (!f) fail("cannot write " + path);
		writeHeader(f, "vectorField", "points", "constant/polyMesh");
		fprintf(f, "%d\n(\n", m.nPoints());
		for (int i = 0; i < m.nPoints(); i++)
			fprintf(f, "(%.17g %.17g %.17g)\n", m.points[(size_t)i * 3], m.points[(size_t)i * 3 + 1], m.points[(size_t)i * 3 + 2]);
		fprintf(f, ")\n");
		fclose(f);
	}
	{
		std::string path = dir + "/faces";
		FILE* f = fopen(path.c_str(), "w");
		if (!f) fail("cannot write " + path);
		writeHeader(f, "faceList", "faces", "constant/polyMesh");
		fprintf(f, "%d\n(\n", m.nFaces());
		for (int i = 0; i < m.nFaces(); i++) {
			int np = m.facePointCount(i);
			fprintf(f, "%d(", np);
			for (int k = 0; k < np; k++) fprintf(f, k ? " %d" : "%d", m.facePoints[(size_t)m.faceOffsets[(size_t)i] + k]);
			fprintf(f, ")\n");
		}
		fprintf(f, ")\n");
		fclose(f);
	}
	writeLabelList(dir + "/owner", "owner", m.owner, note);
	writeLabelList(dir + "/neighbour", "neighbour", m.neighbour, note);
	{
		std::string path = dir + "/boundary";
		FILE* f = fopen(path.c_str(), "w");
		if (!f) fail("cannot write " + path);
		writeHeader(f, "polyBoundaryMesh", "boundary", "constant/polyMesh");
		fprintf(f, "%zu\n(\n", m.patches.size());
		for (const Patch& p : m.patches) {
			fprintf(f, "    %s\n    {\n        type            %s;\n", p.name.c_str(), p.type.c_str());
			if (p.isCyclic()) fprintf(f, "        neighbourPatch  %s;\n", p.neighbourPatch.c_str());
			if (p.isProcessor())
				fprintf(f, "        myProcNo        %d;\n        neighbProcNo    %d;\n", p.myProcNo, p.neighbProcNo);
			if (p.isProcessorCyclic()) fprintf(f, "        referPatch      %s;\n", p.referPatch.c_str());
			fprintf(f, "        nFaces          %d;\n        startFace       %d;\n    }\n", p.nFaces, p.startFace);
		}
		fprintf(f, ")\n");
		fclose(f);
	}
	if (!m.faceProcAddressing.empty()) writeLabelList(dir + "/faceProcAddressing", "faceProcAddressing", m.faceProcAddressing);
	if (!m.cellProcAddressing.empty()) writeLabelList(dir + "/cellProcAddressing", "cellProcAddressing", m.cellProcAddressing);
	if (!m.pointProcAddressing.empty()) writeLabelList(dir + "/pointProcAddressing", "pointProcAddressing", m.pointProcAddressing);
	if (!m.boundaryProcAddressing.empty())
		writeLabelList(dir + "/boundaryProcAddressing", "boundaryProcAddressing", m.boundaryProcAddressing);
	if (!m.cellSubmesh.empty()) writeLabelList(dir + "/cellSubmesh", "cellSubmesh", m.cellSubmesh);
}

// ---------------------------------------------------------------------------------------------
// fields
// ---------------------------------------------------------------------------------------------
std::vector<double> readVolField(const std::string& path, int nCells, int nComp) {
	std::string s = slurp(path);
	const FoamHeader hd = parseHeader(s, path);
	if (hd.binary) {
		// only the contents of a nonuniform list are raw; `uniform` values and everything before the list are text
		size_t pos = s.find("internalField", hd.bodyStart);
		if (pos == std::string::npos) fail("no internalField in " + path);
		const char* end = s.data() + s.size();
		const char* p = s.data() + pos + strlen("internalField");
		skipWsComments(p, end);
		if (strncmp(p, "nonuniform", 10) == 0) {
			p += 10;
			skipWsComments(p, end);
			while (p < end && (unsigned char)*p > ' ' && !(*p >= '0' && *p <= '9')) p++;  // List<scalar>
			if (p < end && *p == '>') p++;
			return readRawScalars(p, end, hd.scalarBytes, nComp, nCells, path);
		}
		size_t semi = s.find(';', pos);   // `uniform X;`: cut the text there and parse it as ascii below
		if (semi == std::string::npos) fail("bad internalField in " + path);
		s.resize(semi + 1);
	}
	stripComments(s);
	skipHeader(s, path);
	size_t pos = s.find("internalField");
	if (pos == std::string::npos) fail("no internalField in " + path);
	const char* end = s.data() + s.size();
	const char* p = s.data() + pos + strlen("internalField");
	skipWs(p, end);
	std::vector<double> v((size_t)nCells * nComp);
	auto readTuple = [&](double* dst) {
		if (nComp == 1) {
			char* e = nullptr;
			dst[0] = strtod(p, &e);
			if (e == p) fail("bad scalar in " + path);
			p = e;
		} else {
			skipWs(p, end);
			if (p >= end || *p != '(') fail("expected '(' in " + path);
			p++;
			for (int k = 0; k < nComp; k++) {
				char* e = nullptr;
				dst[k] = strtod(p, &e);
				if (e == p) fail("bad vector component in " + path);
				p = e;
			}
			skipWs(p, end);
			if (p >= end || *p != ')') fail("expected ')' in " + path);
			p++;
		}
	};
	if (strncmp(p, "uniform", 7) == 0) {
		p += 7;
		double t[3] = {0, 0, 0};
		readTuple(t);
		for (int c = 0; c < nCells; c++)
			for (int k = 0; k < nComp; k++) v[(size_t)c * nComp + k] = t[k];
	} else if (strncmp(p, "nonuniform", 10) == 0) {
		p += 10;
		skipWs(p, end);
		while (p < end && (unsigned char)*p > ' ' && !(*p >= '0' && *p <= '9')) p++;  // List<scalar>
		if (p < end && *p == '>') p++;
		// tolerate "List<scalar>" immediately followed by the size
		long n = readCount(p, end, path);
		if (n != nCells) fail("internalField size mismatch in " + path);
		for (long c = 0; c < n; c++) readTuple(&v[(size_t)c * nComp]);
	} else {
		fail("unsupported internalField in " + path);
	}
	return v;
}

void writeVolField(const std::string& path, const std::string& name, const PolyMesh& m, const std::vector<double>& values,
                   int nComp, int precision, bool binary) {
	// binary: OpenFOAM's binary stream format (controlDict writeFormat binary): the list contents are raw doubles
	FILE* f = fopen(path.c_str(), binary ? "wb" : "w");
	if (!f) fail("cannot write " + path);
	const char* cls = nComp == 1 ? "volScalarField" : "volVectorField";
	if (binary)
		fprintf(f, "FoamFile\n{\n    version     2.0;\n    format      binary;\n    arch        \"LSB;label=32;scalar=64\";\n    class       %s;\n    object      %s;\n}\n\n", cls, name.c_str());
	else
		writeHeader(f, cls, name.c_str(), nullptr);
	fprintf(f, "dimensions      [0 0 0 0 0 0 0];\n\n");
	const size_t n = values.size() / (size_t)nComp;
	if (binary) {
		fprintf(f, "internalField   nonuniform List<%s> \n%zu\n(", nComp == 1 ? "scalar" : "vector", n);
		if (!values.empty() && fwrite(values.data(), sizeof(double), values.size(), f) != values.size()) {
			fclose(f);
			fail("short write on " + path);
		}
		fprintf(f, ")\n;\n\nboundaryField\n{\n");
	} else {
		fprintf(f, "internalField   nonuniform List<%s> \n%zu\n(\n", nComp == 1 ? "scalar" : "vector", n);
		for (size_t c = 0; c < n; c++) {
			if (nComp == 1)
				fprintf(f, "%.*g\n", precision, values[c]);
			else
				fprintf(f, "(%.*g %.*g %.*g)\n", precision, values[c * 3], precision, values[c * 3 + 1], precision, values[c * 3 + 2]);
		}
		fprintf(f, ")\n;\n\nboundaryField\n{\n");
	}
	for (const Patch& p : m.patches) {
		const char* t = "zeroGradient";
		if (p.type == "empty") t = "empty";
		if (p.type == "cyclic") t = "cyclic";
		if (p.type == "processor") t = "processor";
		if (p.type == "processorCyclic") t = "processorCyclic";
		fprintf(f, "    %s\n    {\n        type            %s;\n    }\n", p.name.c_str(), t);
	}
	fprintf(f, "}\n");
	fclose(f);
}

std::string timeName(double t, int precision) {
	std::ostringstream os;
	os.precision(precision);
	os << t;
	return os.str();
}

}  // namespace lfm
