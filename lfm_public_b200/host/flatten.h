// Host-side setup of the hot path: case options, submesh split, per-face geometry, ghost/halo tables and
// initial state, flattened into the pointer-free `lfmgpu_desc` that the CUDA library uploads.
//
// This restates, for a standalone build, what the reference does on the host before the time loop:
//   CInputReader::initialize            src/inputReader.cpp:7-69
//   CPolyMeshReaderOF::initializeSubmesh polyMeshReaderOF/polyMeshReaderOF.cpp:501-610
//   Mesh::readMeshCells / setParams / calculateVNeighbour   src/mesh_reader.cpp:164-615
//   CFDv0_solver::allocate_ghost_cells / assign_pointers / init_params / set_init_conditions /
//                 init_boundary_conditions / reorder_faces   src/cfd_v0.cpp:278-1005
// (In a drop-in build those run unchanged and CFDv0_solver_gpu flattens their result, INTEGRATION.md.)
#pragma once
#include <map>
#include <string>
#include <vector>

#include "foam_io.h"
#include "lfmgpu.h"

namespace lfm {

// Mirror of CInputReader's public members (reference: api/inputReader.h:34-88) + the Foam::Time controls.
struct CaseOptions {
	int commType = 2, haloCommType = 1;
	bool doublePrecision = true;
	bool haveAverage = false, haveForces = false, haveResidual = false;
	int saveForcesStep = 1, printInfoFreq = 1;
	double tStartAverage = 0.0, cflMax = 1.0;
	int solver = 0, dimension = 2, rkOrder = 5;
	bool minmod = false;
	double pInf = 1.0, TInf = 1.0, UInf[3] = {0, 0, 0}, Ls = 0.0, Mach = 0.0, K = 4.34294481903252e-01;
	double Cp = 2.5, molWeight = 11640.3, mu0 = 0.0, Pr = 0.75;
	bool laminar = true;
	double startTime = 0.0, endTime = 0.0, deltaT = 0.0;
	int writeInterval = 1;
	bool adjustTimeStep = false;
	int timePrecision = 6, writePrecision = 6;
};
CaseOptions readCaseOptions(const std::string& caseDir);

struct Fields {
	std::vector<double> p, T, U, alpha;   // [nCells], [nCells], [nCells*3], [nCells] (alpha may be empty when Ls <= 0)
};
// Reads <dir>/<timeName>/{p,T,U[,alpha]} where dir = caseDir or caseDir/processor<rank>.
Fields readFields(const std::string& dir, const CaseOptions& o, int nCells);

class FlatMesh {
public:
	FlatMesh(const PolyMesh& mesh, const Fields& fields, const CaseOptions& opts, int rank, int nRanks);

	int rank() const { return rank_; }
	int neighbourCount() const { return (int)nbrRank_.size(); }
	int neighbourRank(int i) const { return nbrRank_[(size_t)i]; }
	// Setup-time neighbour exchange (replaces the MPI_Isend/Irecv handshakes of mesh_reader.cpp:490-614 and
	// cfd_v0.cpp:646-683): what this rank tells neighbour i, and what neighbour i told this rank.
	std::vector<char> exportFor(int i) const;
	void importFrom(int i, const char* data, size_t bytes);
	// Builds the descriptor; every neighbour must have been imported.
	void finish();

	const lfmgpu_desc& desc() const { return desc_; }
	int nCells() const { return desc_.n_cells; }
	// polyMesh cell index of traversal cell t
	const std::vector<int>& cellGid() const { return cellGid_; }
	const std::vector<std::string>& wallPatchNames() const { return wallPatchNames_; }
	const std::vector<int>& wallPatchIds() const { return wallPatchIds_; }

private:
	template <class P> void buildAll();

	const PolyMesh& m_;
	const Fields& f_;
	CaseOptions o_;
	int rank_, nRanks_;
	int D_;
	bool finished_ = false;

	// submesh bookkeeping
	int nSub_ = 0;
	std::vector<int> cellSub_, subIndex_, travOfCell_, cellGid_;
	std::vector<int> subStart_;
	std::vector<int> subFaceCnt_;
	// per cell valid faces (cells() order, `empty` removed)
	std::vector<int> cvfOff_, cvf_;
	// neighbours
	std::vector<int> nbrRank_;
	std::map<int, int> rank2local_;
	std::vector<std::vector<int>> sendCells_;              // local_cells_to_send (boundary-submesh index)
	std::vector<std::vector<int>> mpiFaces_;               // polyMesh faces towards neighbour i
	struct Remote {
		int ownerBndIndex;
		double x[3];
	};
	std::vector<std::map<std::pair<int, int>, Remote>> remoteByFaceId_;    // imported, keyed (faceId, tag) like faceTagMapping (mesh_reader.cpp:404, 473)
	std::vector<std::vector<int>> recvCells_;              // neigh_cells_to_recv (imported send lists)
	std::vector<bool> imported_;

	// owned storage behind desc_
	std::vector<int32_t> faceOwner_, faceNeigh_, cellSlotFace_, bcCell_, bcKind_, bcPatch_, bcFace_, nbrRank32_, sendStart_,
	    sendCell_, recvStart_;
	std::vector<char> faceS_, faceD_, faceW_, volInv_, sigma_, q0_;
	std::vector<std::string> wallPatchNames_;
	std::vector<int> wallPatchIds_;
	lfmgpu_desc desc_;

	int faceId(int f) const;
	int faceTag(int f) const;
	int cyclicTwin(int f) const;
	int faceNeighbourCell(int f) const;
};

}  // namespace lfm
