// OpenFOAM-free case I/O for the LFM hot path: dictionary, polyMesh, volField parsing and the
// polyMesh geometry (face area/centre, cell centre/volume).
//
// The reference reaches OpenFOAM only through three adapter classes
// (reference: api/dictReaderOF.h, api/polyMeshReaderOF.h, api/runTimeManagerOF.h); everything they
// return is produced here without OpenFOAM.  The geometry follows the published OpenFOAM
// primitiveMesh algorithm (triangle fan about the vertex average for faces, pyramids about the
// face-centre average for cells) -- OpenFOAM v2112 is not vendored in the reference, so this
// part is "parity unpinned" against OpenFOAM itself; the oracle build and the GPU path consume the
// SAME numbers, so hot-path parity does not depend on it.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace lfm {

// ---------------------------------------------------------------------------------------------
// Dictionary
// ---------------------------------------------------------------------------------------------
struct Dict;
struct DictEntry {
	std::string key;
	bool is_dict = false;
	std::shared_ptr<Dict> sub;          // when is_dict
	std::vector<std::string> tokens;    // value tokens (without the trailing ';')
};

struct Dict {
	std::vector<DictEntry> entries;     // insertion order (toc())
	const DictEntry* find(const std::string& key) const;
	// OpenFOAM optionalSubDict: the sub-dictionary if present, else *this
	// (reference: dictReaderOF/dictReaderOF.cpp:110-117)
	const Dict& optionalSubDict(const std::string& key) const;
	bool has(const std::string& key) const { return find(key) != nullptr; }
	double scalar(const std::string& key) const;                 // hard-fails (throws) when missing
	double scalarOr(const std::string& key, double def) const;
	bool boolean(const std::string& key) const;
	bool booleanOr(const std::string& key, bool def) const;
	std::string word(const std::string& key) const;
	std::string wordOr(const std::string& key, const std::string& def) const;
	std::string valueString(const std::string& key) const;       // tokens joined by ' '
};

// Parses an OpenFOAM ASCII dictionary file (FoamFile header kept as the sub-dict "FoamFile").
Dict parseDictFile(const std::string& path);
Dict parseDictString(const std::string& text);
bool parseSwitch(const std::string& tok, bool& out);

// ---------------------------------------------------------------------------------------------
// polyMesh
// ---------------------------------------------------------------------------------------------
struct Patch {
	std::string name;
	std::string type;            // wall / patch / empty / cyclic / processor / ...
	int nFaces = 0;
	int startFace = 0;
	std::string neighbourPatch;  // cyclic
	int neighbPatchID = -1;      // resolved index of neighbourPatch
	int myProcNo = -1;           // processor
	int neighbProcNo = -1;       // processor
	std::string referPatch;      // processorCyclic: the cyclic patch whose faces these were before the decomposition
	int referPatchID = -1;       // resolved index of referPatch
	int tag = -1;                // processorCyclic: message tag if the file gives one
	bool isProcessorCyclic() const { return type == "processorCyclic"; }
	bool isProcessor() const { return type == "processor" || isProcessorCyclic(); }   // processorCyclicPolyPatch derives from processorPolyPatch
	bool isCyclic() const { return type == "cyclic"; }
	bool coupled() const { return isProcessor() || isCyclic(); }
};

struct PolyMesh {
	// topology
	std::vector<double> points;        // [nPoints*3]
	std::vector<int> faceOffsets;      // [nFaces+1] into facePoints
	std::vector<int> facePoints;
	std::vector<int> owner;            // [nFaces]
	std::vector<int> neighbour;        // [nInternalFaces]
	std::vector<Patch> patches;
	int nCells = 0;
	// optional addressing (processorN meshes); empty when absent
	std::vector<int> faceProcAddressing, cellProcAddressing, pointProcAddressing, boundaryProcAddressing;
	std::vector<int> cellSubmesh;      // optional polyMesh/cellSubmesh override
	// derived: cells()[c] face order = owned faces ascending, then neighbour faces ascending
	std::vector<int> cellFaceOffsets;  // [nCells+1]
	std::vector<int> cellFaces;
	// geometry
	std::vector<double> faceAreas;     // [nFaces*3] (owner-outward)
	std::vector<double> faceCentres;   // [nFaces*3]
	std::vector<double> cellCentres;   // [nCells*3]
	std::vector<double> cellVolumes;   // [nCells]

	int nPoints() const { return (int)(points.size() / 3); }
	int nFaces() const { return (int)owner.size(); }
	int nInternalFaces() const { return (int)neighbour.size(); }
	int nNonProcessor() const;          // index of the first processor patch
	int whichPatch(int face) const;     // -1 for internal faces
	int processorTag(int patch) const;   // processorPolyPatch::tag() * (2 owner() - 1); 0 for other patches (polyMeshReaderOF.cpp:460-475)
	int facePointCount(int f) const { return faceOffsets[f + 1] - faceOffsets[f]; }

	void buildCells();                  // cellFaceOffsets / cellFaces
	void computeGeometry();             // faceAreas ... cellVolumes
	void resolvePatches();              // neighbPatchID
	void finalize() { resolvePatches(); buildCells(); computeGeometry(); }
};

// Reads <meshDir>/{points,faces,owner,neighbour,boundary} (+ optional *ProcAddressing, cellSubmesh).
// ASCII only; throws std::runtime_error on malformed input or "format binary".
PolyMesh readPolyMesh(const std::string& meshDir);
void writePolyMesh(const PolyMesh& m, const std::string& meshDir);

// ---------------------------------------------------------------------------------------------
// vol fields (internalField only; boundaryField is carried through verbatim on write)
// ---------------------------------------------------------------------------------------------
// nComp = 1 (volScalarField) or 3 (volVectorField).  Expands "uniform" to nCells values.
std::vector<double> readVolField(const std::string& path, int nCells, int nComp);
// Writes a field with zeroGradient patches (empty patches get "empty", processor/cyclic their own type).
void writeVolField(const std::string& path, const std::string& name, const PolyMesh& m,
                   const std::vector<double>& values, int nComp, int precision = 17, bool binary = false);

std::string timeName(double t, int precision = 12);   // OpenFOAM "general" time formatting

// hpath.cpp: restatement of the reference's hpathRenumber plugin; order[new] = old, stats as in lfmhost_hpath_order
std::vector<int> hpathOrder(const PolyMesh& mesh, double stats[4]);

}  // namespace lfm
