// Host-side flattening of an LFM case into lfmgpu_desc (see flatten.h for the reference functions restated).
// Arithmetic that defines the hot path's inputs (S, d, weight_linear, vol_inv, sigma, q0, gas constants) is
// written with the reference's expression trees and evaluated in PRECISION (template parameter P) so that
// the descriptor is bit-identical to what CFDv0_solver<P,D,F> holds after reorder_faces().
#include "flatten.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace lfm {

namespace {
[[noreturn]] void fail(const std::string& msg) { throw std::runtime_error("lfm::flatten: " + msg); }
const int GHOST_SM = 10000;   // MAX_MPI_RANKS (reference: api/mpi_env.h:4)

template <class T> void put(std::vector<char>& buf, const T& v) {
	const char* p = (const char*)&v;
	buf.insert(buf.end(), p, p + sizeof(T));
}
template <class T> T get(const char*& p, const char* end) {
	if (p + sizeof(T) > end) fail("truncated neighbour exchange buffer");
	T v;
	memcpy(&v, p, sizeof(T));
	p += sizeof(T);
	return v;
}
}  // namespace

// -------------------------------------------------------------------------------------------------
// options  (reference: src/inputReader.cpp:13-64; Foam::Time controls read by runTimeManagerOF)
// -------------------------------------------------------------------------------------------------
CaseOptions readCaseOptions(const std::string& caseDir) {
	CaseOptions o;
	Dict cd = parseDictFile(caseDir + "/system/controlDict");
	const Dict& l = cd.optionalSubDict("lfm");
	const Dict& post = l.optionalSubDict("post");
	o.commType = (int)l.scalarOr("commType", 2);
	o.haloCommType = (int)l.scalarOr("haloCommType", 1);
	o.doublePrecision = l.booleanOr("doublePrecision", true);
	o.haveAverage = post.booleanOr("haveAverage", false);
	o.tStartAverage = post.scalarOr("tStartAverage", 0);
	o.haveForces = post.booleanOr("haveForces", false);
	o.haveResidual = post.booleanOr("haveResiduals", false);
	o.saveForcesStep = (int)post.scalarOr("saveForcesStep", 1);
	o.printInfoFreq = (int)post.scalarOr("printInfoFreq", 1);
	o.cflMax = cd.scalarOr("maxCo", 1);
	o.startTime = cd.scalarOr("startTime", 0.0);
	o.endTime = cd.scalar("endTime");
	o.deltaT = cd.scalar("deltaT");
	o.writeInterval = (int)cd.scalarOr("writeInterval", 1);
	o.adjustTimeStep = cd.booleanOr("adjustTimeStep", false);
	o.timePrecision = (int)cd.scalarOr("timePrecision", 6);
	o.writePrecision = (int)cd.scalarOr("writePrecision", 6);

	Dict fs = parseDictFile(caseDir + "/system/fvSchemes");
	const Dict& fl = fs.optionalSubDict("lfm");
	o.solver = (int)fl.scalarOr("solver", 0);
	o.dimension = (int)fl.scalarOr("dimension", 2);
	o.rkOrder = (int)fl.scalarOr("rkOrder", 5);
	o.minmod = fl.booleanOr("minmodExists", false);

	Dict sp = parseDictFile(caseDir + "/constant/spongeDict");
	o.pInf = sp.scalar("pinf");
	o.TInf = sp.scalar("Tinf");
	o.UInf[0] = sp.scalarOr("Uinf_x", 0);
	o.UInf[1] = sp.scalarOr("Uinf_y", 0);
	o.UInf[2] = sp.scalarOr("Uinf_z", 0);
	o.Ls = sp.scalar("Ls");
	o.Mach = sp.scalar("M");
	o.K = sp.scalarOr("k", 4.34294481903252e-01);

	Dict th = parseDictFile(caseDir + "/constant/thermophysicalProperties");
	const Dict& mix = th.optionalSubDict("mixture");
	o.Cp = mix.optionalSubDict("thermodynamics").scalar("Cp");
	o.molWeight = mix.optionalSubDict("specie").scalar("molWeight");
	o.mu0 = mix.optionalSubDict("transport").scalar("mu");
	o.Pr = mix.optionalSubDict("transport").scalar("Pr");

	Dict tp = parseDictFile(caseDir + "/constant/turbulenceProperties");
	o.laminar = (tp.valueString("simulationType") == "laminar");
	return o;
}

Fields readFields(const std::string& dir, const CaseOptions& o, int nCells) {
	Fields f;
	const std::string t = dir + "/" + timeName(o.startTime, o.timePrecision) + "/";
	f.p = readVolField(t + "p", nCells, 1);
	f.T = readVolField(t + "T", nCells, 1);
	f.U = readVolField(t + "U", nCells, 3);
	if (o.Ls > 0) f.alpha = readVolField(t + "alpha", nCells, 1);
	return f;
}

// -------------------------------------------------------------------------------------------------
int FlatMesh::cyclicTwin(int f) const {
	// reference: polyMeshReaderOF.cpp:226-248 getCyclicFaceIndex
	const int b = m_.whichPatch(f);
	if (b == -1 || !m_.patches[(size_t)b].coupled()) return f;
	const int nb = m_.patches[(size_t)b].neighbPatchID;
	if (nb == -1) return f;
	return f - m_.patches[(size_t)b].startFace + m_.patches[(size_t)nb].startFace;
}

int FlatMesh::faceNeighbourCell(int f) const {
	// reference: polyMeshReaderOF.cpp:149-162 getFaceNeighbour
	if (f < m_.nInternalFaces()) return m_.neighbour[(size_t)f];
	const int c = cyclicTwin(f);
	return (c == f) ? -1 : m_.owner[(size_t)c];
}

int FlatMesh::faceTag(int f) const {
	// reference: polyMeshReaderOF.cpp:460-475 getBoundaryTag of the face's patch
	const int b = m_.whichPatch(f);
	return b < 0 ? 0 : m_.processorTag(b);
}

int FlatMesh::faceId(int f) const {
	// reference: polyMeshReaderOF.cpp:251-261 getFaceId: the global face (faceProcAddressing) behind plain processor
	// patches (tag +-1), the index inside the patch behind processorCyclic patches (any other tag)
	const int tag = faceTag(f);
	if (tag < -1 || tag > 1) return f - m_.patches[(size_t)m_.whichPatch(f)].startFace;
	if ((size_t)f < m_.faceProcAddressing.size()) return std::abs(m_.faceProcAddressing[(size_t)f]);
	return f;
}

FlatMesh::FlatMesh(const PolyMesh& mesh, const Fields& fields, const CaseOptions& opts, int rank, int nRanks)
    : m_(mesh), f_(fields), o_(opts), rank_(rank), nRanks_(nRanks), D_(opts.dimension) {
	memset(&desc_, 0, sizeof desc_);
	if (D_ != 2 && D_ != 3) fail("invalid dimension");
	const int nc = m_.nCells;
	// ---- submesh split (polyMeshReaderOF.cpp:501-551) -------------------------------------------------
	cellSub_.assign((size_t)nc, 1);
	if ((int)m_.cellSubmesh.size() == nc) {
		cellSub_ = m_.cellSubmesh;
	} else {
		std::vector<char> interiorPoint((size_t)m_.nPoints(), 1);
		for (const Patch& p : m_.patches) {
			if (p.type == "empty") continue;
			for (int f = p.startFace; f < p.startFace + p.nFaces; f++)
				for (int k = m_.faceOffsets[(size_t)f]; k < m_.faceOffsets[(size_t)f + 1]; k++) interiorPoint[(size_t)m_.facePoints[(size_t)k]] = 0;
		}
		for (int c = 0; c < nc; c++) {
			bool interior = true;
			for (int i = m_.cellFaceOffsets[(size_t)c]; i < m_.cellFaceOffsets[(size_t)c + 1] && interior; i++) {
				const int f = m_.cellFaces[(size_t)i];
				for (int k = m_.faceOffsets[(size_t)f]; k < m_.faceOffsets[(size_t)f + 1] && interior; k++)
					interior = interiorPoint[(size_t)m_.facePoints[(size_t)k]] != 0;
			}
			cellSub_[(size_t)c] = interior ? 1 : 0;
		}
	}
	nSub_ = 0;
	for (int c = 0; c < nc; c++) nSub_ = std::max(nSub_, cellSub_[(size_t)c] + 1);
	if (nSub_ > LFMGPU_MAX_SUBMESH) fail("too many submeshes");
	std::vector<int> cnt((size_t)nSub_, 0);
	subIndex_.assign((size_t)nc, 0);
	for (int c = 0; c < nc; c++) subIndex_[(size_t)c] = cnt[(size_t)cellSub_[(size_t)c]]++;
	subStart_.assign((size_t)nSub_ + 1, 0);
	for (int s = 0; s < nSub_; s++) subStart_[(size_t)s + 1] = subStart_[(size_t)s] + cnt[(size_t)s];
	travOfCell_.assign((size_t)nc, 0);
	cellGid_.assign((size_t)nc, 0);
	for (int c = 0; c < nc; c++) {
		travOfCell_[(size_t)c] = subStart_[(size_t)cellSub_[(size_t)c]] + subIndex_[(size_t)c];
		cellGid_[(size_t)travOfCell_[(size_t)c]] = c;
	}
	// ---- valid faces per cell (polyMeshReaderOF.cpp:326-364) ----------------------------------------------
	std::vector<char> validFace((size_t)m_.nFaces(), 1);
	for (const Patch& p : m_.patches)
		if (p.type == "empty")
			for (int f = p.startFace; f < p.startFace + p.nFaces; f++) validFace[(size_t)f] = 0;
	cvfOff_.assign((size_t)nc + 1, 0);
	for (int c = 0; c < nc; c++) {
		int n = 0;
		for (int i = m_.cellFaceOffsets[(size_t)c]; i < m_.cellFaceOffsets[(size_t)c + 1]; i++) n += validFace[(size_t)m_.cellFaces[(size_t)i]];
		cvfOff_[(size_t)c + 1] = cvfOff_[(size_t)c] + n;
	}
	cvf_.resize((size_t)cvfOff_[(size_t)nc]);
	for (int c = 0; c < nc; c++) {
		int o = cvfOff_[(size_t)c];
		for (int i = m_.cellFaceOffsets[(size_t)c]; i < m_.cellFaceOffsets[(size_t)c + 1]; i++)
			if (validFace[(size_t)m_.cellFaces[(size_t)i]]) cvf_[(size_t)o++] = m_.cellFaces[(size_t)i];
	}
	subFaceCnt_.assign((size_t)nSub_, 0);
	for (int c = 0; c < nc; c++) subFaceCnt_[(size_t)cellSub_[(size_t)c]] = std::max(subFaceCnt_[(size_t)cellSub_[(size_t)c]], cvfOff_[(size_t)c + 1] - cvfOff_[(size_t)c]);
	for (int c = 0; c < nc; c++)
		if (cvfOff_[(size_t)c + 1] - cvfOff_[(size_t)c] != subFaceCnt_[(size_t)cellSub_[(size_t)c]])
			fail("non-uniform face count inside a submesh (the reference dereferences a null neighbour there, cfd_v0.cpp:301-310)");
	// ---- MPI neighbours and send lists (cfd_v0.cpp:436-463, 616-644), boundary submesh only ---------------
	const int nNonProc = m_.nNonProcessor();
	std::vector<std::vector<char>> recorded;
	for (int c = 0; c < nc; c++) {
		if (cellSub_[(size_t)c] != 0) continue;
		for (int i = cvfOff_[(size_t)c]; i < cvfOff_[(size_t)c + 1]; i++) {
			const int f = cvf_[(size_t)i];
			if (f < m_.nInternalFaces()) continue;
			const int b = m_.whichPatch(f);
			if (b < nNonProc) continue;
			const int r = m_.patches[(size_t)b].neighbProcNo;
			auto it = rank2local_.find(r);
			int li;
			if (it == rank2local_.end()) {
				li = (int)nbrRank_.size();
				rank2local_[r] = li;
				nbrRank_.push_back(r);
				sendCells_.emplace_back();
				mpiFaces_.emplace_back();
				recorded.emplace_back((size_t)subStart_[1], 0);
			} else {
				li = it->second;
			}
			mpiFaces_[(size_t)li].push_back(f);
			const int bi = subIndex_[(size_t)c];
			if (!recorded[(size_t)li][(size_t)bi]) {
				recorded[(size_t)li][(size_t)bi] = 1;
				sendCells_[(size_t)li].push_back(bi);
			}
		}
	}
	for (int c = 0; c < nc; c++) {
		if (cellSub_[(size_t)c] == 0) continue;
		for (int i = cvfOff_[(size_t)c]; i < cvfOff_[(size_t)c + 1]; i++) {
			const int f = cvf_[(size_t)i];
			if (f >= m_.nInternalFaces() && faceNeighbourCell(f) == -1)
				fail("an interior-submesh cell has a boundary face (cellSubmesh override inconsistent)");
		}
	}
	remoteByFaceId_.resize(nbrRank_.size());
	recvCells_.resize(nbrRank_.size());
	imported_.assign(nbrRank_.size(), false);
}

std::vector<char> FlatMesh::exportFor(int i) const {
	// payload of mesh_reader.cpp:527-556 (t_cell_center: faceId, tag, owner's boundary-submesh index, xc - xf)
	// followed by local_cells_to_send[i] (cfd_v0.cpp:667-683)
	std::vector<char> buf;
	const std::vector<int>& faces = mpiFaces_[(size_t)i];
	put<int32_t>(buf, (int32_t)faces.size());
	put<int32_t>(buf, (int32_t)sendCells_[(size_t)i].size());
	for (int f : faces) {
		const int c = m_.owner[(size_t)f];
		put<int32_t>(buf, (int32_t)faceId(f));
		put<int32_t>(buf, (int32_t)faceTag(f));
		put<int32_t>(buf, (int32_t)subIndex_[(size_t)c]);
		for (int k = 0; k < 3; k++) put<double>(buf, m_.cellCentres[(size_t)c * 3 + k] - m_.faceCentres[(size_t)f * 3 + k]);
	}
	for (int s : sendCells_[(size_t)i]) put<int32_t>(buf, (int32_t)s);
	return buf;
}

void FlatMesh::importFrom(int i, const char* data, size_t bytes) {
	const char* p = data;
	const char* end = data + bytes;
	const int nf = get<int32_t>(p, end);
	const int ns = get<int32_t>(p, end);
	if (nf != (int)mpiFaces_[(size_t)i].size()) fail("neighbour reports a different number of shared faces");
	remoteByFaceId_[(size_t)i].clear();
	for (int k = 0; k < nf; k++) {
		const int id = get<int32_t>(p, end);
		const int tag = get<int32_t>(p, end);
		Remote r;
		r.ownerBndIndex = get<int32_t>(p, end);
		for (int d = 0; d < 3; d++) r.x[d] = get<double>(p, end);
		remoteByFaceId_[(size_t)i][std::make_pair(id, tag)] = r;
	}
	recvCells_[(size_t)i].resize((size_t)ns);
	for (int k = 0; k < ns; k++) recvCells_[(size_t)i][(size_t)k] = get<int32_t>(p, end);
	imported_[(size_t)i] = true;
}

void FlatMesh::finish() {
	for (size_t i = 0; i < imported_.size(); i++)
		if (!imported_[i]) fail("neighbour " + std::to_string(nbrRank_[i]) + " was not imported");
	if (o_.doublePrecision)
		buildAll<double>();
	else
		buildAll<float>();
	finished_ = true;
}

template <class P>
void FlatMesh::buildAll() {
	const int D = D_;
	const int nc = m_.nCells;
	const int nBnd = subStart_[1];
	int F = 0;
	for (int s = 0; s < nSub_; s++) F = std::max(F, subFaceCnt_[(size_t)s]);

	// ---- physical boundary ghosts: boundaries[patch] = {cell, slot} (cfd_v0.cpp:382-431) ----------------
	const int nPatch = (int)m_.patches.size();
	const int nNonProc = m_.nNonProcessor();
	std::vector<int> facePatch((size_t)m_.nFaces(), -1);
	for (int b = 0; b < nPatch; b++)
		for (int f = m_.patches[(size_t)b].startFace; f < m_.patches[(size_t)b].startFace + m_.patches[(size_t)b].nFaces; f++) facePatch[(size_t)f] = b;
	std::vector<std::vector<std::pair<int, int>>> boundaries((size_t)nPatch);
	std::vector<int> bcLocalOfSlot((size_t)cvfOff_[(size_t)nc], -1);
	for (int t = 0; t < nBnd; t++) {
		const int c = cellGid_[(size_t)t];
		for (int i = cvfOff_[(size_t)c]; i < cvfOff_[(size_t)c + 1]; i++) {
			const int f = cvf_[(size_t)i];
			if (f < m_.nInternalFaces()) continue;
			const int b = facePatch[(size_t)f];
			if (b >= nNonProc) continue;                 // CELL_MPI
			if (faceNeighbourCell(f) != -1) continue;    // cyclic -> ordinary neighbour
			bcLocalOfSlot[(size_t)i] = (int)boundaries[(size_t)b].size();
			boundaries[(size_t)b].push_back({t, i - cvfOff_[(size_t)c]});
		}
	}
	std::vector<int> bcBase((size_t)nPatch + 1, 0);
	for (int b = 0; b < nPatch; b++) bcBase[(size_t)b + 1] = bcBase[(size_t)b] + (int)boundaries[(size_t)b].size();
	const int nBc = bcBase[(size_t)nPatch];
	// roles (cfd_v0.cpp:973-1005)
	std::vector<int> patchKind((size_t)nPatch, LFMGPU_BC_NONE);
	wallPatchIds_.clear();
	wallPatchNames_.clear();
	for (int b = 0; b < nPatch; b++) {
		const Patch& p = m_.patches[(size_t)b];
		if (p.type == "wall") {
			patchKind[(size_t)b] = LFMGPU_BC_WALL;
			wallPatchIds_.push_back(b);
			wallPatchNames_.push_back(p.name);
		}
		// `type farfield` (cfd_v0.cpp:986-988, 1136-1215) picks the inlet or the outlet state per face from the sign of
		// Udirection . S with Udirection = (cos, sin)(m_dAoA): the reference never assigns m_dAoA (api/cfdv0_solver.h:141 is its only
		// other mention), so its farfield ghosts depend on an uninitialised value.  There is no defined behaviour to reproduce:
		// refuse the case instead of running it with ghosts that were never set.
		if (p.type == "farfield") fail("patch '" + p.name + "' has type farfield: not served (the reference's farfield state reads the unset m_dAoA)");
		if (p.type == "patch") {
			if (p.name == "inlet" || p.name == "Inlet" || p.name == "Inflow" || p.name == "inflow") patchKind[(size_t)b] = LFMGPU_BC_INLET;
			if (p.name == "outlet" || p.name == "Outlet" || p.name == "Outflow" || p.name == "outflow") patchKind[(size_t)b] = LFMGPU_BC_OUTLET;
		}
	}
	// ---- MPI ghosts ------------------------------------------------------------------------------------------
	const int nNbr = (int)nbrRank_.size();
	recvStart_.assign((size_t)nNbr + 1, 0);
	sendStart_.assign((size_t)nNbr + 1, 0);
	std::vector<std::map<int, int>> recvPos((size_t)nNbr);
	for (int n = 0; n < nNbr; n++) {
		recvStart_[(size_t)n + 1] = recvStart_[(size_t)n] + (int)recvCells_[(size_t)n].size();
		sendStart_[(size_t)n + 1] = sendStart_[(size_t)n] + (int)sendCells_[(size_t)n].size();
		for (size_t i = 0; i < recvCells_[(size_t)n].size(); i++) recvPos[(size_t)n][recvCells_[(size_t)n][i]] = (int)i;
	}
	const int nMpi = recvStart_[(size_t)nNbr];

	// ---- per (cell, slot) data in the ORIGINAL slot order ----------------------------------------------------
	struct Slot {
		int polyFace;
		int nbFlat;          // flat state index of the neighbour
		int keySm, keyId;    // (sm_id, id) of the neighbour for reorder_faces
		P S[3], d[3], w;
	};
	std::vector<Slot> slots((size_t)cvfOff_[(size_t)nc]);
	int ghostIdRunning = nBnd;   // ghost_bnd ids continue after the boundary submesh (cfd_v0.cpp:421-431)
	std::vector<int> bcGhostId((size_t)nBc, 0);
	for (int b = 0; b < nPatch; b++)
		for (size_t k = 0; k < boundaries[(size_t)b].size(); k++) bcGhostId[(size_t)bcBase[(size_t)b] + k] = ghostIdRunning++;

	for (int c = 0; c < nc; c++) {
		for (int i = cvfOff_[(size_t)c]; i < cvfOff_[(size_t)c + 1]; i++) {
			Slot& s = slots[(size_t)i];
			const int f = cvf_[(size_t)i];
			s.polyFace = f;
			const int own = m_.owner[(size_t)f];
			const int other = faceNeighbourCell(f);
			const int nb = (own == c) ? other : own;
			const double vsign = (own == c) ? 1.0 : -1.0;
			const double* xc = &m_.cellCentres[(size_t)c * 3];
			const double* xf = &m_.faceCentres[(size_t)f * 3];
			// init_params (cfd_v0.cpp:779-861)
			P dP[3], dN[3];
			for (int k = 0; k < D; k++) s.S[k] = vsign * m_.faceAreas[(size_t)f * 3 + k];
			for (int k = 0; k < D; k++) dP[k] = xc[k] - xf[k];
			bool isNone = false;
			double vneigh[3] = {0, 0, 0};   // calculateVNeighbour (mesh_reader.cpp:404-613), a property of the face
			if (nb != -1) {
				const double* xo = &m_.cellCentres[(size_t)own * 3];
				const double* xn = &m_.cellCentres[(size_t)other * 3];
				for (int k = 0; k < D; k++) vneigh[k] = xn[k] - xo[k];
				const int twin = cyclicTwin(f);
				if (twin != f) {
					const double* xt = &m_.faceCentres[(size_t)twin * 3];
					for (int k = 0; k < D; k++) vneigh[k] += xf[k] - xt[k];
				}
				s.nbFlat = travOfCell_[(size_t)nb];
				s.keySm = cellSub_[(size_t)nb];
				s.keyId = subIndex_[(size_t)nb];
			} else {
				const int b = facePatch[(size_t)f];
				if (b < nNonProc) {
					isNone = true;
					const int g = bcBase[(size_t)b] + bcLocalOfSlot[(size_t)i];
					s.nbFlat = nc + g;
					s.keySm = GHOST_SM;
					s.keyId = bcGhostId[(size_t)g];
				} else {
					const int n = rank2local_.at(m_.patches[(size_t)b].neighbProcNo);
					auto it = remoteByFaceId_[(size_t)n].find(std::make_pair(faceId(f), -faceTag(f)));   // the sender's tag is the reverse of ours (mesh_reader.cpp:460-473)
					if (it == remoteByFaceId_[(size_t)n].end()) fail("processor face without a partner on the neighbour rank");
					const Remote& r = it->second;
					for (int k = 0; k < D; k++) vneigh[k] = r.x[k] + xf[k] - xc[k];
					auto rp = recvPos[(size_t)n].find(r.ownerBndIndex);
					if (rp == recvPos[(size_t)n].end()) fail("neighbour cell missing from its send list");
					s.nbFlat = nc + nBc + recvStart_[(size_t)n] + rp->second;
					s.keySm = GHOST_SM;
					s.keyId = r.ownerBndIndex;
				}
			}
			if (isNone) {
				P S_mag_sqrt = 0.0;
				P dP_dot_S = 0.0;
				for (int k = 0; k < D; k++) S_mag_sqrt += s.S[k] * s.S[k];
				for (int k = 0; k < D; k++) dP_dot_S += dP[k] * s.S[k];
				for (int k = 0; k < D; k++) s.d[k] = 2.0 * std::abs(dP_dot_S) * s.S[k] / S_mag_sqrt;
			} else {
				for (int k = 0; k < D; k++) s.d[k] = vsign * vneigh[k];
			}
			for (int k = 0; k < D; k++) dN[k] = dP[k] + s.d[k];
			s.w = 0.0;
			P divider = 0.0;
			for (int k = 0; k < D; k++) {
				s.w += s.S[k] * dN[k];
				divider += (-s.S[k] * dP[k] + s.S[k] * dN[k]);
			}
			s.w /= divider;
		}
	}

	// ---- reorder_faces (cfd_v0.cpp:278-339) + face list in traversal order ----------------------------------
	cellSlotFace_.assign((size_t)nc * F, 0);
	std::vector<int> lfmFaceOfPoly((size_t)m_.nFaces(), -1);
	std::vector<int> slotOfOrder((size_t)nc * F, -1);   // reordered slot -> index into `slots`
	faceOwner_.clear();
	faceNeigh_.clear();
	std::vector<P> fS, fd, fw;
	std::vector<int> subFaceStart((size_t)nSub_ + 1, 0);
	for (int t = 0; t < nc; t++) {
		const int c = cellGid_[(size_t)t];
		const int n = cvfOff_[(size_t)c + 1] - cvfOff_[(size_t)c];
		int nValid = 0, nInvalid = F - 1;
		const int csm = cellSub_[(size_t)c], cid = subIndex_[(size_t)c];
		for (int k = 0; k < n; k++) {
			const Slot& s = slots[(size_t)cvfOff_[(size_t)c] + k];
			bool valid = true;
			if (s.keySm < csm)
				valid = false;
			else if (s.keySm == csm && s.keyId < cid)
				valid = false;
			const int dst = valid ? nValid++ : nInvalid--;
			slotOfOrder[(size_t)t * F + dst] = cvfOff_[(size_t)c] + k;
		}
		for (int k = 0; k < nValid; k++) {
			const Slot& s = slots[(size_t)slotOfOrder[(size_t)t * F + k]];
			const int id = (int)faceOwner_.size();
			faceOwner_.push_back(t);
			faceNeigh_.push_back(s.nbFlat);
			for (int d = 0; d < D; d++) {
				fS.push_back(s.S[d]);
				fd.push_back(s.d[d]);
			}
			fw.push_back(s.w);
			lfmFaceOfPoly[(size_t)s.polyFace] = id;
			cellSlotFace_[(size_t)t * F + k] = id + 1;
		}
		subFaceStart[(size_t)cellSub_[(size_t)c] + 1] = (int)faceOwner_.size();
	}
	for (int s = 0; s < nSub_; s++)
		if (subFaceStart[(size_t)s + 1] < subFaceStart[(size_t)s]) subFaceStart[(size_t)s + 1] = subFaceStart[(size_t)s];
	for (int t = 0; t < nc; t++)
		for (int k = 0; k < F; k++) {
			if (cellSlotFace_[(size_t)t * F + k] != 0) continue;
			const int si = slotOfOrder[(size_t)t * F + k];
			if (si < 0) continue;
			const Slot& s = slots[(size_t)si];
			const int seenByOwner = cyclicTwin(s.polyFace);   // the neighbour reaches this connection through the twin
			const int id = lfmFaceOfPoly[(size_t)seenByOwner];
			if (id < 0) fail("internal error: invalid slot without an owning face");
			cellSlotFace_[(size_t)t * F + k] = -(id + 1);
		}
	const int nFacesFlat = (int)faceOwner_.size();

	// ---- boundary tables ----------------------------------------------------------------------------------------
	bcCell_.assign((size_t)nBc, 0);
	bcKind_.assign((size_t)nBc, 0);
	bcPatch_.assign((size_t)nBc, 0);
	bcFace_.assign((size_t)nBc, -1);
	for (int b = 0; b < nPatch; b++)
		for (size_t k = 0; k < boundaries[(size_t)b].size(); k++) {
			const int g = bcBase[(size_t)b] + (int)k;
			bcCell_[(size_t)g] = boundaries[(size_t)b][k].first;
			bcKind_[(size_t)g] = patchKind[(size_t)b];
			bcPatch_[(size_t)g] = b;
		}
	for (int fidx = 0; fidx < nFacesFlat; fidx++) {
		const int nb = faceNeigh_[(size_t)fidx];
		if (nb >= nc && nb < nc + nBc) bcFace_[(size_t)(nb - nc)] = fidx;
	}

	// ---- cell data, initial conditions, constants (cfd_v0.cpp:859-965) ---------------------------------------
	const P ZERO = 0.0, ONE = 1.0;
	(void)ZERO;
	std::vector<P> volInv((size_t)nc), sigma((size_t)nc, (P)0.0), q0((size_t)nc * (D + 2));
	const P m_dRuniversal = 8.31447;
	P m_dpInf = o_.pInf, m_dTInf = o_.TInf, m_dUInf[3] = {(P)o_.UInf[0], (P)o_.UInf[1], (P)o_.UInf[2]};
	P m_dMu0 = o_.mu0, m_dMolWeight = o_.molWeight, m_dCp = o_.Cp;
	const P dLs = o_.Ls;
	const P dPrandtl = o_.Pr;
	P m_dRgas = m_dRuniversal / m_dMolWeight * 1000;
	const P dCv = m_dCp - m_dRgas;
	P m_dGamma = m_dCp / dCv;
	P m_drhoInf = m_dpInf / (m_dRgas * m_dTInf);
	P m_dGammaMinusOne = m_dGamma - ONE;
	P dUMagInf2 = 0;
	for (int k = 0; k < D; k++) dUMagInf2 += m_dUInf[k] * m_dUInf[k];
	P m_dEInf = m_dpInf / (m_drhoInf * m_dGammaMinusOne) + 0.5 * dUMagInf2;
	P m_dRgas_inv = ONE / m_dRgas;
	P m_dPr_inv = ONE / dPrandtl;
	P dSigma0 = 0;
	if (dLs > 0) {
		const P dK = o_.K;
		const P dMach = o_.Mach;
		dSigma0 = (3.0 * (1.0 - dMach * dMach) / (dK * dLs));
		if ((int)f_.alpha.size() != nc) fail("alpha field required when Ls > 0");
	}
	for (int t = 0; t < nc; t++) {
		const int c = cellGid_[(size_t)t];
		volInv[(size_t)t] = ONE / m_.cellVolumes[(size_t)c];
		if (dLs > 0) {
			const double dDist = f_.alpha[(size_t)c];
			if (dDist < dLs) {
				const double r = (dLs - dDist) / dLs;
				sigma[(size_t)t] = dSigma0 * (r * r);   // pow(x, 2.0): gcc expands to x*x
			}
		}
		const double* dUVW = &f_.U[(size_t)c * 3];
		const P dPressure = f_.p[(size_t)c];
		const P dTemperature = f_.T[(size_t)c];
		P dUMag2 = 0;
		for (int k = 0; k < D; k++) dUMag2 += dUVW[k] * dUVW[k];
		const P dRho = dPressure / (m_dRgas * dTemperature);
		const P dEnergy = dPressure / (dRho * m_dGammaMinusOne) + 0.5 * dUMag2;
		q0[(size_t)t * (D + 2)] = dRho;
		for (int k = 0; k < D; k++) q0[(size_t)t * (D + 2) + k + 1] = dRho * dUVW[k];
		q0[(size_t)t * (D + 2) + D + 1] = dRho * dEnergy;
	}

	// ---- publish ---------------------------------------------------------------------------------------------------
	auto store = [](std::vector<char>& dst, const std::vector<P>& src) {
		dst.resize(src.size() * sizeof(P));
		if (!src.empty()) memcpy(dst.data(), src.data(), dst.size());
	};
	store(faceS_, fS);
	store(faceD_, fd);
	store(faceW_, fw);
	store(volInv_, volInv);
	store(sigma_, sigma);
	store(q0_, q0);
	nbrRank32_.assign(nbrRank_.begin(), nbrRank_.end());
	sendCell_.clear();
	for (int n = 0; n < nNbr; n++) sendCell_.insert(sendCell_.end(), sendCells_[(size_t)n].begin(), sendCells_[(size_t)n].end());

	lfmgpu_desc& d = desc_;
	memset(&d, 0, sizeof d);
	d.precision = (int)sizeof(P);
	d.dim = D;
	d.max_slots = F;
	d.n_sub = nSub_;
	for (int s = 0; s <= nSub_; s++) {
		d.sub_cell_start[s] = subStart_[(size_t)s];
		d.sub_face_start[s] = subFaceStart[(size_t)s];
	}
	for (int s = 0; s < nSub_; s++) d.sub_face_cnt[s] = subFaceCnt_[(size_t)s];
	d.n_cells = nc;
	d.n_faces = nFacesFlat;
	d.n_bc_ghosts = nBc;
	d.n_mpi_ghosts = nMpi;
	d.face_owner = faceOwner_.data();
	d.face_neigh = faceNeigh_.data();
	d.face_S = faceS_.data();
	d.face_d = faceD_.data();
	d.face_w = faceW_.data();
	d.vol_inv = volInv_.data();
	d.sponge_sigma = sigma_.data();
	d.q0 = q0_.data();
	d.cell_gid = cellGid_.data();
	d.cell_slot_face = cellSlotFace_.data();
	d.bc_cell = bcCell_.data();
	d.bc_kind = bcKind_.data();
	d.bc_patch = bcPatch_.data();
	d.bc_face = bcFace_.data();
	d.n_nbr = nNbr;
	d.nbr_rank = nbrRank32_.data();
	d.send_start = sendStart_.data();
	d.send_cell = sendCell_.data();
	d.recv_start = recvStart_.data();
	lfmgpu_consts& c = d.c;
	c.gamma = m_dGamma;
	c.gamma_m1 = m_dGammaMinusOne;
	c.Rgas_inv = m_dRgas_inv;
	c.mu = m_dMu0;
	c.Cp = m_dCp;
	c.Pr_inv = m_dPr_inv;
	c.rhoInf = m_drhoInf;
	for (int k = 0; k < 3; k++) c.UInf[k] = m_dUInf[k];
	c.EInf = m_dEInf;
	c.pInf = m_dpInf;
	c.TInf = m_dTInf;
	// Runge-Kutta coefficients (cfd_v0.cpp:86-128), stored in PRECISION
	std::vector<P> Ak, Bk;
	switch (o_.rkOrder) {
		case 0:
		case 1: Ak = {(P)0.0}; Bk = {(P)1.0}; break;
		case 2: Ak = {(P)0.0, (P)-1.0}; Bk = {(P)1.0, (P)0.5}; break;
		case 3:
		case 4:
			Ak = {(P)0.0, (P)(-756391.0 / 934407.0), (P)(-36441873.0 / 15625000.0), (P)(-1953125.0 / 1085297.0)};
			Bk = {(P)(8.0 / 141.0), (P)(6627.0 / 2000.0), (P)(609375.0 / 1085297.0), (P)(198961.0 / 526383.0)};
			break;
		case 5:
			Ak = {(P)0.0, (P)-0.4178904745, (P)-1.192151694643, (P)-1.697784692471, (P)-1.514183444257};
			Bk = {(P)0.1496590219993, (P)0.3792103129999, (P)0.8229550293869, (P)0.6994504559488, (P)0.1530572479681};
			break;
		default: fail("unsupported rkOrder");
	}
	c.rk_order = o_.rkOrder;   // the stage loop runs rkOrder times (mesh_solver.cpp:500)
	if ((int)Ak.size() < o_.rkOrder) fail("rkOrder exceeds the coefficient table (the reference would read out of bounds)");
	for (size_t k = 0; k < Ak.size(); k++) {
		c.Ak[k] = Ak[k];
		c.Bk[k] = Bk[k];
	}
	c.comm_type = o_.commType;
}

}  // namespace lfm
