// Hamiltonian-path cell numbering: a restatement of the reference's renumberMesh plugin
// (hpathRenumber/hpathRenumber.C, method `hpath` of renumberMeshDict; OpenFOAM plugin, 2D meshes) on the
// OpenFOAM-free PolyMesh, so that config 2 of BASELINE.json ("2D unstructured with hpath renumbering") can be
// prepared without OpenFOAM.  Result: order[new] = old with the boundary submesh first (a walk that always moves
// to the nearest unwalked boundary cell), then the interior submesh (a walk that hugs the rim of what is left,
// best of several starts and of two next-cell rules), cells that no walk reached ("dead ends") appended to
// their part.  PINNED against the plugin itself: oracle/_ref/hpath_plugin is hpathRenumber.C compiled unchanged against a
// stand-in for the few OpenFOAM types it touches (oracle/shim/openfoam_stub/), and tests/test_hpath_vs_plugin.py requires
// the two orders to be equal on the zoo's meshes, on larger shuffled meshes and on the reference's own example meshes
// (525 000 quads, 1 959 342 triangle prisms); tests/test_renumber.py keeps the invariants.
//
// Decisions of the plugin that shape the numbering and are kept, quirks included:
//   * `empty` patches do not count as boundary (hpathRenumber.C:137-146);
//   * submesh 0 = cells with a point on a non-empty boundary face (:381-423) -- the reference's own split;
//   * the boundary walk looks for the next cell with a breadth-first search over ALL cells, at most 10 rings
//     deep, in cells() face order (:228-266); a neighbour that would be cut off is set aside, not walked (:283-303);
//   * interior starts: rim cells passing three local tests, sorted by angle about the mean cell centre and
//     thinned to about 2 000 000 / nCells trials (:440-601);
//   * next-cell rules (:636-843): "standard" = rim cells, then cells next to the rim, then (once neither exists)
//     the free neighbour farthest from the mesh centre; "spiral" = rim cell, else the candidate whose rim /
//     walked neighbour was walked longest ago; a walk that gets stuck steps back up to 60 cells (:928-973);
//   * the marking loop of :868-874 flags the OWNER of a face between two submeshes whichever side is interior.
// Deviations (the plugin's behaviour is undefined there): a missing neighbour (-1) is skipped where the plugin
// indexes a vector with it (:521-524); the trial budget is clamped to >= 1 (the plugin divides by
// 2 000 000 / nCells, which is zero above two million cells, :577-579).
#include <algorithm>
#include <cmath>
#include <deque>
#include <vector>

#include "foam_io.h"

namespace lfm {

namespace {

struct Walker {
	const PolyMesh& m;
	std::vector<char> validFace;
	std::vector<int> sub;          // submesh of each cell: 0 boundary, 1 interior
	double centre[3] = {0, 0, 0};  // mean of the cell centres (getCenter, :425-430)

	explicit Walker(const PolyMesh& mesh) : m(mesh) {}

	int nCells() const { return m.nCells; }
	int nbr(int c, int f) const {   // getNei (:181-186)
		if (f >= m.nInternalFaces()) return -1;
		const int o = m.owner[(size_t)f];
		return o == c ? m.neighbour[(size_t)f] : o;
	}
	const int* facesBegin(int c) const { return m.cellFaces.data() + m.cellFaceOffsets[(size_t)c]; }
	const int* facesEnd(int c) const { return m.cellFaces.data() + m.cellFaceOffsets[(size_t)c + 1]; }
	double dist2ToCentre(int c) const {
		double s = 0;
		for (int k = 0; k < 3; k++) {
			const double d = m.cellCentres[(size_t)c * 3 + k] - centre[k];
			s += d * d;
		}
		return s;
	}

	void prepare() {
		validFace.assign((size_t)m.nFaces(), 1);
		for (const Patch& p : m.patches)
			if (p.type == "empty")
				for (int f = p.startFace; f < p.startFace + p.nFaces; f++) validFace[(size_t)f] = 0;
		// submesh split (:381-423)
		std::vector<char> rimPoint(m.points.size() / 3, 0);
		sub.assign((size_t)nCells(), 1);
		for (int f = m.nInternalFaces(); f < m.nFaces(); f++) {
			if (!validFace[(size_t)f]) continue;
			for (int k = m.faceOffsets[(size_t)f]; k < m.faceOffsets[(size_t)f + 1]; k++) rimPoint[(size_t)m.facePoints[(size_t)k]] = 1;
			sub[(size_t)m.owner[(size_t)f]] = 0;
		}
		for (int f = 0; f < m.nInternalFaces(); f++)
			for (int k = m.faceOffsets[(size_t)f]; k < m.faceOffsets[(size_t)f + 1]; k++)
				if (rimPoint[(size_t)m.facePoints[(size_t)k]]) {
					sub[(size_t)m.owner[(size_t)f]] = 0;
					sub[(size_t)m.neighbour[(size_t)f]] = 0;
					break;
				}
		const int n = nCells();
		for (int c = 0; c < n; c++)
			for (int k = 0; k < 3; k++) centre[k] += m.cellCentres[(size_t)c * 3 + k];
		for (int k = 0; k < 3; k++) centre[k] /= n;
	}

	// a neighbour of `from` that nobody but `from` can still reach inside from's submesh (:267-276)
	bool cutOff(int from, int cand, const std::vector<char>& walked) const {
		for (const int* f = facesBegin(cand); f != facesEnd(cand); ++f) {
			const int x = nbr(cand, *f);
			if (x < 0 || x == from || walked[(size_t)x] || sub[(size_t)x] != sub[(size_t)from]) continue;
			return false;
		}
		return true;
	}

	// ---------------------------------------------------------------- boundary submesh (:189-378) ----
	// nearest (in face hops through any cell, at most 10) unwalked cell of submesh `part`
	int nearestUnwalked(int from, int part, int skip, const std::vector<char>& walked, std::vector<char>& seen, std::vector<int>& touched) const {
		for (int x : touched) seen[(size_t)x] = 0;
		touched.clear();
		std::deque<int> q;
		q.push_back(from);
		q.push_back(-1);
		int depth = 0;
		while (!q.empty() && depth <= 10) {
			const int c = q.front();
			q.pop_front();
			if (c == -1) {   // ring finished
				depth++;
				q.push_back(-1);
				continue;
			}
			if (seen[(size_t)c]) continue;
			seen[(size_t)c] = 1;
			touched.push_back(c);
			if (c != from && !walked[(size_t)c] && c != skip && sub[(size_t)c] == part) return c;
			for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
				const int x = nbr(c, *f);
				if (x > -1 && !seen[(size_t)x]) q.push_back(x);
			}
		}
		return -1;
	}

	int cutOffNeighbourBnd(int c, int last, const std::vector<char>& walked) const {   // :278-303
		int left = (int)(facesEnd(c) - facesBegin(c));
		for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
			const int x = nbr(c, *f);
			if (x < 0 || walked[(size_t)x] || sub[(size_t)x] != sub[(size_t)c]) --left;
		}
		if (left <= 1) return -1;
		for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
			const int x = nbr(c, *f);
			if (x < 0 || walked[(size_t)x] || x == last) continue;
			if (sub[(size_t)x] != sub[(size_t)c]) continue;
			if (cutOff(c, x, walked)) return x;
		}
		return -1;
	}

	bool walkBoundaryFrom(int start, int part, std::vector<int>& out) const {   // findBoundaryHpath (:305-378)
		const int n = nCells();
		std::vector<char> walked((size_t)n, 0), setAside((size_t)n, 0), seen((size_t)n, 0);
		std::vector<int> touched, path, aside;
		const int firstNext = nearestUnwalked(start, part, -1, walked, seen, touched);
		int nValid = 0, last = -1;
		for (const int* f = facesBegin(start); f != facesEnd(start); ++f) nValid += validFace[(size_t)*f] ? 1 : 0;
		for (const int* f = facesBegin(start); f != facesEnd(start); ++f) {
			if (*f >= m.nInternalFaces()) continue;
			const int x = nbr(start, *f);
			if (sub[(size_t)x] != part) continue;
			if (nValid == 3) {
				if (x != firstNext) last = x;
			} else {
				last = x;
			}
		}
		int next = start;
		do {
			const int c = next;
			walked[(size_t)c] = 1;
			path.push_back(c);
			const int cut = cutOffNeighbourBnd(c, last, walked);
			if (cut > -1) {
				aside.push_back(cut);
				setAside[(size_t)cut] = 1;
			}
			next = nearestUnwalked(c, part, cut, walked, seen, touched);
		} while (next != -1);
		for (int c = 0; c < n; c++)
			if (sub[(size_t)c] == part && !walked[(size_t)c] && !setAside[(size_t)c]) aside.push_back(c);
		out = path;
		for (int c : aside)
			if (!walked[(size_t)c]) out.push_back(c);
		return (long)out.size() == std::count(sub.begin(), sub.end(), 0);
	}

	bool boundaryOrder(int part, std::vector<int>& out) const {   // getBoundaryHpath (:188-226)
		const int n = nCells(), kStarts = 10;
		std::vector<int> starts;
		for (int c = 0; c < n; c++)
			if (sub[(size_t)c] == part) {
				starts.push_back(c);
				break;
			}
		for (int c = 0; c < n && (int)starts.size() < kStarts; c++) {
			if (sub[(size_t)c] != part) continue;
			int nValid = 0, good = 0;
			for (const int* f = facesBegin(c); f != facesEnd(c); ++f) nValid += validFace[(size_t)*f] ? 1 : 0;
			if (nValid < 4) continue;
			for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
				if (!validFace[(size_t)*f]) continue;
				if (*f < m.nInternalFaces() && sub[(size_t)nbr(c, *f)] == part && ++good > 1) break;
			}
			if (good > 1) starts.push_back(c);
		}
		for (int s : starts)
			if (walkBoundaryFrom(s, part, out)) return true;
		return false;
	}

	// ---------------------------------------------------------------- interior submesh (:440-1076) ----
	std::vector<int> interiorStarts(int part) const {   // getInteriorStartingCells (:440-601)
		const int n = nCells();
		std::vector<int> rim((size_t)n, 0), rimCells;
		for (int f = 0; f < m.nInternalFaces(); f++) {
			const int a = m.owner[(size_t)f], b = m.neighbour[(size_t)f];
			if (sub[(size_t)a] != sub[(size_t)b]) rim[(size_t)a] = rim[(size_t)b] = 1;
		}
		for (int c = 0; c < n; c++) {
			if (sub[(size_t)c] != part) rim[(size_t)c] = 0;
			if (rim[(size_t)c]) rimCells.push_back(c);
		}
		if (rimCells.empty()) return {};
		// the only same-part neighbour of a rim cell with fewer than two of them
		std::vector<char> nextToLonely((size_t)n, 0);
		for (int c : rimCells) {
			int cnt = 0, only = -1;
			for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
				if (!validFace[(size_t)*f]) continue;
				const int x = nbr(c, *f);
				if (x < 0 || sub[(size_t)x] != part) continue;
				cnt++;
				only = x;
			}
			if (cnt < 2 && only >= 0) nextToLonely[(size_t)only] = 1;
		}
		std::vector<int> cand;
		for (int c : rimCells) {
			bool ok = true;
			for (const int* f = facesBegin(c); f != facesEnd(c) && ok; ++f) {
				if (!validFace[(size_t)*f]) continue;
				const int x = nbr(c, *f);
				if (x < 0 || sub[(size_t)x] != part) continue;
				if (nextToLonely[(size_t)x]) ok = false;
			}
			if (!ok) continue;
			int roomy = 0;      // neighbours with at least two other same-part neighbours
			for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
				if (!validFace[(size_t)*f]) continue;
				const int x = nbr(c, *f);
				if (x < 0 || sub[(size_t)x] != part) continue;
				int others = 0;
				for (const int* g = facesBegin(x); g != facesEnd(x); ++g) {
					if (!validFace[(size_t)*g]) continue;
					const int y = nbr(x, *g);
					if (y < 0 || sub[(size_t)y] != part || y == c) continue;
					others++;
				}
				if (others > 1) roomy++;
			}
			if (roomy < 2) continue;
			int nearRim = 0;    // neighbours that touch a rim cell
			for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
				const int x = nbr(c, *f);
				if (x < 0 || sub[(size_t)x] != part) continue;
				bool touches = false;
				for (const int* g = facesBegin(x); g != facesEnd(x) && !touches; ++g) {
					const int y = nbr(x, *g);
					if (y < 0) continue;
					if (rim[(size_t)y]) touches = true;
				}
				if (touches) nearRim++;
			}
			if (nearRim >= 2) cand.push_back(c);
		}
		if (cand.empty()) {    // the rim cell farthest from the centre
			int far = rimCells[0];
			double best = dist2ToCentre(far);
			for (int c : rimCells) {
				const double d = dist2ToCentre(c);
				if (d > best) {
					best = d;
					far = c;
				}
			}
			cand.push_back(far);
		}
		// spread the trials around the rim: by angle about the centre (single precision, as stored by the plugin)
		std::vector<std::pair<float, float>> byAngle(cand.size());
		for (size_t i = 0; i < cand.size(); i++) {
			const int c = cand[i];
			byAngle[i].first = (float)std::atan2(m.cellCentres[(size_t)c * 3 + 1] - centre[1], m.cellCentres[(size_t)c * 3] - centre[0]);
			byAngle[i].second = (float)i;
		}
		std::sort(byAngle.begin(), byAngle.end());
		const int cnt = (int)cand.size();
		const int budget = std::max(1, 2000000 / std::max(1, n));
		const int every = budget > cnt ? 1 : std::max(1, cnt / budget);
		std::vector<int> starts;
		for (int i = 0; i < cnt; i += every) starts.push_back(cand[(size_t)byAngle[(size_t)i].second]);
		return starts;
	}

	int cutOffNeighbourInt(int c, const std::vector<char>& walked) const {   // :604-633
		int left = 0;
		for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
			if (!validFace[(size_t)*f]) continue;
			const int x = nbr(c, *f);
			if (x < 0 || sub[(size_t)x] != sub[(size_t)c] || walked[(size_t)x]) continue;
			left++;
		}
		if (left <= 1) return -1;
		for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
			if (!validFace[(size_t)*f]) continue;
			const int x = nbr(c, *f);
			if (x < 0 || sub[(size_t)x] != sub[(size_t)c] || walked[(size_t)x]) continue;
			if (cutOff(c, x, walked)) return x;
		}
		return -1;
	}

	int nextFarthest(int c, int skip, const std::vector<char>& walked) const {   // :636-668
		int freeCnt = 0, single = -1, best = -1;
		double bestD = 0;
		for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
			const int x = nbr(c, *f);
			if (x < 0 || sub[(size_t)x] != sub[(size_t)c] || walked[(size_t)x]) continue;
			freeCnt++;
			single = x;
			if (x == skip) continue;
			const double d = std::sqrt(dist2ToCentre(x));
			if (bestD < d) {
				bestD = d;
				best = x;
			}
		}
		if (freeCnt == 1) return single;
		if (freeCnt == 0) return -1;
		return best;
	}

	int nextHugRim(int c, int skip, const std::vector<char>& walked, const std::vector<char>& onRim, const std::vector<char>& byRim, bool& rimExhausted) const {   // :671-722
		const int part = sub[(size_t)c];
		auto open = [&](int x) { return x >= 0 && sub[(size_t)x] == part && !walked[(size_t)x] && x != skip; };
		for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
			const int x = nbr(c, *f);
			if (open(x) && onRim[(size_t)x]) return x;
		}
		for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
			const int x = nbr(c, *f);
			if (open(x) && byRim[(size_t)x]) return x;
		}
		for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
			const int x = nbr(c, *f);
			if (!open(x)) continue;
			for (const int* g = facesBegin(x); g != facesEnd(x); ++g) {
				const int y = nbr(x, *g);
				if (y < 0 || sub[(size_t)y] != part || walked[(size_t)y]) continue;
				if (byRim[(size_t)y]) return y;
			}
		}
		for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
			const int x = nbr(c, *f);
			if (!open(x)) continue;
			rimExhausted = true;
			return x;
		}
		return -1;
	}

	int nextSpiral(int c, int skip, const std::vector<char>& walked, const std::vector<char>& onRim, const std::vector<char>& byRim, const std::vector<int>& when) const {   // :724-843
		const int part = sub[(size_t)c];
		const int never = 2 * nCells();
		std::vector<int> rimNext, plain;
		for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
			const int x = nbr(c, *f);
			if (x < 0 || sub[(size_t)x] != part || walked[(size_t)x] || x == skip) continue;
			if (onRim[(size_t)x]) return x;
			(byRim[(size_t)x] ? rimNext : plain).push_back(x);
		}
		if (!rimNext.empty()) {
			int oldest = never, pick = -1;
			for (int x : rimNext)
				for (const int* g = facesBegin(x); g != facesEnd(x); ++g) {
					const int y = nbr(x, *g);
					if (y < 0 || sub[(size_t)y] != part || !onRim[(size_t)y]) continue;
					if (oldest > when[(size_t)y]) {
						oldest = when[(size_t)y];
						pick = x;
					}
				}
			if (pick != -1) return pick;
		}
		if (plain.empty()) return -1;
		if (plain.size() == 1) return plain[0];
		int oldest = never, pick = -1;
		for (int x : plain)
			for (const int* g = facesBegin(x); g != facesEnd(x); ++g) {
				const int y = nbr(x, *g);
				if (y < 0 || !walked[(size_t)y] || y == c) continue;
				if (oldest > when[(size_t)y]) {
					oldest = when[(size_t)y];
					pick = x;
				}
			}
		if (pick != -1) return pick;
		for (int pass = 0; pass < 2; pass++) {   // two rings out; the first pass ignores what was walked in the last three steps
			oldest = never;
			pick = -1;
			for (int x : plain)
				for (const int* g = facesBegin(x); g != facesEnd(x); ++g) {
					const int y = nbr(x, *g);
					if (y < 0 || (pass == 0 && y == c)) continue;
					for (const int* h = facesBegin(y); h != facesEnd(y); ++h) {
						const int z = nbr(y, *h);
						if (z < 0 || !walked[(size_t)z]) continue;
						if (pass == 0 && when[(size_t)c] - when[(size_t)z] < 3) continue;
						if (oldest > when[(size_t)z]) {
							oldest = when[(size_t)z];
							pick = x;
						}
					}
				}
			if (pick > -1) return pick;
		}
		return plain[0];
	}

	// one interior walk (findInteriorHpathNoDeadend, :876-1026); rule 0 = standard, 1 = spiral
	bool walkInteriorFrom(int start, int part, int rule, std::vector<int>& path, std::vector<int>& leftOver) const {
		const int n = nCells();
		std::vector<char> onRim((size_t)n, 0), byRim((size_t)n, 0), walked((size_t)n, 0);
		const long total = std::count(sub.begin(), sub.end(), part);
		for (int f = 0; f < m.nInternalFaces(); f++) {
			const int a = m.owner[(size_t)f], b = m.neighbour[(size_t)f];
			if (sub[(size_t)a] != sub[(size_t)b] && (sub[(size_t)a] == part || sub[(size_t)b] == part)) onRim[(size_t)a] = 1;   // the owner, whichever side (:868-874)
		}
		for (int f = 0; f < m.nInternalFaces(); f++) {
			const int a = m.owner[(size_t)f], b = m.neighbour[(size_t)f];
			if (sub[(size_t)a] != sub[(size_t)b] || sub[(size_t)a] != part) continue;
			if (onRim[(size_t)a]) byRim[(size_t)b] = 1;
			if (onRim[(size_t)b]) byRim[(size_t)a] = 1;
		}
		std::vector<int> seq((size_t)n, -1), when((size_t)n, -1), aside;
		int len = 0, asideCnt = 0, next = start;
		bool rimExhausted = false;
		do {
			const int c = next;
			when[(size_t)c] = len;
			seq[(size_t)len++] = c;
			walked[(size_t)c] = 1;
			const int cut = cutOffNeighbourInt(c, walked);
			if (cut > -1) {
				asideCnt++;
				aside.push_back(cut);
			}
			if (rule == 1)
				next = nextSpiral(c, cut, walked, onRim, byRim, when);
			else if (rimExhausted)
				next = nextFarthest(c, cut, walked);
			else
				next = nextHugRim(c, cut, walked, onRim, byRim, rimExhausted);
			if (next == -1 && len + asideCnt < total) {   // stuck: step back until another way out appears (:928-973)
				int back = 1, exit = -1;
				for (;;) {
					if (len - back <= 0) break;
					const int blocked = seq[(size_t)(len - back)];
					const int from = seq[(size_t)(len - back - 1)];
					if (from == -1) break;
					exit = nextFarthest(from, blocked, walked);
					if (exit != -1) break;
					++back;
					if (back == 60) break;
				}
				if (exit == -1) break;
				for (int i = 1; i < back - 1; i++) {
					const int x = seq[(size_t)(len - i)];
					walked[(size_t)x] = 0;
					aside.push_back(x);
					when[(size_t)x] = -1;
					seq[(size_t)(len - i)] = -1;
					++asideCnt;
				}
				next = exit;
				len -= back;
			}
			if (next != -1) walked[(size_t)next] = 1;
		} while (next != -1);
		path.assign(seq.begin(), seq.begin() + len);
		// everything of the part the walk did not reach, in the order the path passes by it (:982-1012)
		std::vector<char> reached((size_t)n, 0), noted((size_t)n, 0);
		for (int c : path) reached[(size_t)c] = 1;
		leftOver.clear();
		for (int c : path)
			for (const int* f = facesBegin(c); f != facesEnd(c); ++f) {
				const int x = nbr(c, *f);
				if (x < 0 || sub[(size_t)x] != part) continue;
				if (!reached[(size_t)x] && !noted[(size_t)x]) {
					leftOver.push_back(x);
					noted[(size_t)x] = 1;
				}
			}
		for (int c = 0; c < n; c++)
			if (sub[(size_t)c] == part && !reached[(size_t)c] && !noted[(size_t)c]) {
				leftOver.push_back(c);
				noted[(size_t)c] = 1;
			}
		return total == (long)(path.size() + leftOver.size());
	}

	bool interiorOrder(int part, std::vector<int>& path, std::vector<int>& leftOver, double& pathFraction) const {   // getInteriorHpath (:1029-1076)
		const std::vector<int> starts = interiorStarts(part);
		const long total = std::count(sub.begin(), sub.end(), part);
		int bestLen = -1, bestStart = -1, bestRule = 0;
		for (int rule = 0; rule <= 1; rule++)
			for (size_t t = 0; t < starts.size(); t++) {
				if (!walkInteriorFrom(starts[t], part, rule, path, leftOver)) continue;
				if (bestLen < (int)path.size()) {
					bestLen = (int)path.size();
					bestStart = (int)t;
					bestRule = rule;
				}
			}
		if (bestStart < 0) return false;
		walkInteriorFrom(starts[(size_t)bestStart], part, bestRule, path, leftOver);
		pathFraction = total ? (double)path.size() / (double)total : 1.0;
		return true;
	}
};

}  // namespace

// order[new] = old.  stats: [0] boundary-submesh cells, [1] 1 if the boundary walk succeeded, [2] 1 if an interior walk
// succeeded, [3] fraction of the interior submesh on the path itself (the plugin's "Best hpath percent" / 100).
std::vector<int> hpathOrder(const PolyMesh& mesh, double stats[4]) {
	Walker w(mesh);
	w.prepare();
	const int n = mesh.nCells;
	std::vector<int> order;
	order.reserve((size_t)n);
	std::vector<int> part0;
	const bool okB = w.boundaryOrder(0, part0);
	if (okB)
		order = part0;
	else
		for (int c = 0; c < n; c++)
			if (w.sub[(size_t)c] == 0) order.push_back(c);   // identity inside the part (:159-164)
	const size_t nBnd = order.size();
	std::vector<int> path, leftOver;
	double frac = 0.0;
	const bool okI = w.interiorOrder(1, path, leftOver, frac);
	if (okI) {
		order.insert(order.end(), path.begin(), path.end());
		order.insert(order.end(), leftOver.begin(), leftOver.end());
	} else {
		for (int c = 0; c < n; c++)
			if (w.sub[(size_t)c] == 1) order.push_back(c);
	}
	if (stats) {
		stats[0] = (double)nBnd;
		stats[1] = okB ? 1.0 : 0.0;
		stats[2] = okI ? 1.0 : 0.0;
		stats[3] = frac;
	}
	return order;
}

}  // namespace lfm
