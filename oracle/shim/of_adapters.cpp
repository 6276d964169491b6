// OpenFOAM-free implementations of the reference's three adapter classes, with the exact class names and
// headers of /root/reference/api (CDictReaderOF, CPolyMeshReaderOF, CRunTimeManagerOF), so that the
// reference's src/*.cpp + info/lfm_solve.cpp link UNCHANGED into oracle/_ref/lfm_solve_ref.
//
// TEST INFRASTRUCTURE ONLY (oracle build).  Behaviour restated from
//   reference: dictReaderOF/dictReaderOF.cpp, polyMeshReaderOF/polyMeshReaderOF.cpp,
//              runTimeManagerOF/runTimeManagerOF.cpp   
// on top of lfm_public_b200/host/foam_io (the product's own case reader, so oracle and GPU path see
// bit-identical geometry and fields).
#include <sys/stat.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

#include "api/dictReaderOF.h"
#include "api/polyMeshReaderOF.h"
#include "api/runTimeManagerOF.h"
#include "foam_io.h"

namespace {

struct ShimField {
	std::string name;
	int nComp;
	std::vector<double> v;
};

struct ShimMesh {
	lfm::PolyMesh mesh;
	std::vector<ShimField*> fields;   // everything AUTO_WRITE
};

struct ShimTime {
	std::string caseDir;              // "." or "processorN"
	double startTime = 0, endTime = 0, deltaT = 0, value = 0;
	int timeIndex = 0;
	int writeInterval = 1;
	bool adjust = false;
	int timePrecision = 12, writePrecision = 17;
	bool writeBinary = false;      // controlDict writeFormat binary (the reference's 3D cases)
	ShimMesh* mesh = nullptr;
	std::string timeName() const { return lfm::timeName(value, timePrecision); }
};

[[noreturn]] void fatal(const std::string& msg) {
	std::cerr << "--> FOAM FATAL ERROR (shim): " << msg << std::endl;
	exit(1);
}

const lfm::Dict* descend(const lfm::Dict* d, const std::vector<std::string>& sub) {
	for (const std::string& s : sub) d = &d->optionalSubDict(s);
	return d;
}

}  // namespace

// =================================================================================================
// CDictReaderOF  (reference: dictReaderOF/dictReaderOF.cpp)
// =================================================================================================
CDictReaderOF::CDictReaderOF(std::string dictName, std::string dictDirectory) {
	try {
		m_pIODict = new lfm::Dict(lfm::parseDictFile("./" + dictDirectory + "/" + dictName));
	} catch (const std::exception& e) {
		fatal(e.what());
	}
}

CDictReaderOF::~CDictReaderOF() {
	delete static_cast<lfm::Dict*>(m_pIODict);
	m_pIODict = nullptr;
}

const void* CDictReaderOF::getDict(std::vector<std::string> subDictList) {
	return descend(static_cast<const lfm::Dict*>(m_pIODict), subDictList);
}

std::vector<std::string> CDictReaderOF::extractSubDictList(std::string subDict) {
	std::vector<std::string> out;
	if (subDict.empty()) return out;
	size_t prev = 0, pos;
	while ((pos = subDict.find('/', prev)) != std::string::npos) {
		out.push_back(subDict.substr(prev, pos - prev));
		prev = pos + 1;
	}
	out.push_back(subDict.substr(prev));
	return out;
}

std::string CDictReaderOF::readDictWord(std::string paramName, std::vector<std::string> subDictList, const std::string sDefault,
                                        const bool bAllowDefault) {
	const lfm::Dict* d = static_cast<const lfm::Dict*>(getDict(subDictList));
	try {
		return bAllowDefault ? d->wordOr(paramName, sDefault) : d->word(paramName);
	} catch (const std::exception& e) {
		fatal(e.what());
	}
}

std::string CDictReaderOF::readDictFilename(std::string paramName, std::vector<std::string> subDictList, const std::string sDefault,
                                            const bool bAllowDefault) {
	return readDictWord(paramName, subDictList, sDefault, bAllowDefault);
}

bool CDictReaderOF::readDictBool(std::string paramName, std::vector<std::string> subDictList, const bool bDefault,
                                 const bool bAllowDefault) {
	const lfm::Dict* d = static_cast<const lfm::Dict*>(getDict(subDictList));
	try {
		return bAllowDefault ? d->booleanOr(paramName, bDefault) : d->boolean(paramName);
	} catch (const std::exception& e) {
		fatal(e.what());
	}
}

double CDictReaderOF::readDictScalar(std::string paramName, std::vector<std::string> subDictList, const double dDefault,
                                     const bool bAllowDefault) {
	const lfm::Dict* d = static_cast<const lfm::Dict*>(getDict(subDictList));
	try {
		return bAllowDefault ? d->scalarOr(paramName, dDefault) : d->scalar(paramName);
	} catch (const std::exception& e) {
		fatal(e.what());
	}
}

std::string CDictReaderOF::readDictString(std::string paramName, std::vector<std::string> subDictList, const std::string,
                                          const bool) {
	const lfm::Dict* d = static_cast<const lfm::Dict*>(getDict(subDictList));
	try {
		return d->valueString(paramName);
	} catch (const std::exception& e) {
		fatal(e.what());
	}
}

std::vector<double> CDictReaderOF::readDictVector(std::string paramName, std::vector<std::string> subDictList) {
	const lfm::Dict* d = static_cast<const lfm::Dict*>(getDict(subDictList));
	const lfm::DictEntry* e = d->find(paramName);
	if (!e || e->tokens.size() < 5) fatal("keyword " + paramName + " is not a vector");
	return {atof(e->tokens[1].c_str()), atof(e->tokens[2].c_str()), atof(e->tokens[3].c_str())};
}

std::vector<std::string> CDictReaderOF::getKeyOrDictList(const bool bKey, std::vector<std::string> subDictList) {
	const lfm::Dict* d = static_cast<const lfm::Dict*>(getDict(subDictList));
	std::vector<std::string> out;
	for (const lfm::DictEntry& e : d->entries)
		if (e.is_dict != bKey) out.push_back(e.key);
	return out;
}

// =================================================================================================
// CRunTimeManagerOF  (reference: runTimeManagerOF/runTimeManagerOF.cpp; Foam::Time semantics per SURVEY App. C)
// =================================================================================================
CRunTimeManagerOF::CRunTimeManagerOF(const int nRank) {
	ShimTime* t = new ShimTime;
	t->caseDir = ".";
	if (nRank >= 0) t->caseDir = "processor" + std::to_string(nRank);
	try {
		lfm::Dict cd = lfm::parseDictFile("./system/controlDict");
		t->startTime = cd.scalarOr("startTime", 0.0);
		t->endTime = cd.scalar("endTime");
		t->deltaT = cd.scalar("deltaT");
		t->writeInterval = (int)cd.scalarOr("writeInterval", 1);
		if (cd.wordOr("writeControl", "timeStep") != "timeStep") fatal("shim supports writeControl timeStep only");
		if (cd.wordOr("startFrom", "startTime") != "startTime") fatal("shim supports startFrom startTime only");
		t->adjust = cd.booleanOr("adjustTimeStep", false);
		t->timePrecision = (int)cd.scalarOr("timePrecision", 6);
		const char* wp = getenv("LFM_WRITE_PRECISION");
		t->writePrecision = wp ? atoi(wp) : (int)cd.scalarOr("writePrecision", 6);
		t->writeBinary = cd.wordOr("writeFormat", "ascii") == "binary";
	} catch (const std::exception& e) {
		fatal(e.what());
	}
	t->value = t->startTime;
	m_pRunTime = t;
}

CRunTimeManagerOF::~CRunTimeManagerOF() {
	delete static_cast<ShimTime*>(m_pRunTime);
}

double CRunTimeManagerOF::getStartTime() const { return static_cast<ShimTime*>(m_pRunTime)->startTime; }
double CRunTimeManagerOF::getEndTime() const { return static_cast<ShimTime*>(m_pRunTime)->endTime; }
double CRunTimeManagerOF::getCurrentTime() const { return static_cast<ShimTime*>(m_pRunTime)->value; }
double CRunTimeManagerOF::getDeltaTime() const { return static_cast<ShimTime*>(m_pRunTime)->deltaT; }
void CRunTimeManagerOF::setDeltaTime(const double dDeltaTime) { static_cast<ShimTime*>(m_pRunTime)->deltaT = dDeltaTime; }
bool CRunTimeManagerOF::isAdjustDeltaTime() const { return static_cast<ShimTime*>(m_pRunTime)->adjust; }
int CRunTimeManagerOF::getTimeStep() const { return static_cast<ShimTime*>(m_pRunTime)->timeIndex; }

void CRunTimeManagerOF::advanceTime() {
	ShimTime* t = static_cast<ShimTime*>(m_pRunTime);
	t->value = t->value + t->deltaT;
	t->timeIndex++;
}

bool CRunTimeManagerOF::isRunning() {
	ShimTime* t = static_cast<ShimTime*>(m_pRunTime);
	return t->value < (t->endTime - 0.5 * t->deltaT);
}

bool CRunTimeManagerOF::isWriteTime() {
	ShimTime* t = static_cast<ShimTime*>(m_pRunTime);
	return t->writeInterval > 0 && (t->timeIndex % t->writeInterval) == 0;
}

void CRunTimeManagerOF::writeResults() {
	ShimTime* t = static_cast<ShimTime*>(m_pRunTime);
	if (!t->mesh) return;
	std::string dir = t->caseDir + "/" + t->timeName();
	mkdir(dir.c_str(), 0777);
	for (ShimField* f : t->mesh->fields)
		lfm::writeVolField(dir + "/" + f->name, f->name, t->mesh->mesh, f->v, f->nComp, t->writePrecision, t->writeBinary);
}

// =================================================================================================
// CPolyMeshReaderOF  (reference: polyMeshReaderOF/polyMeshReaderOF.cpp)
// =================================================================================================
#define MESH (static_cast<ShimMesh*>(m_pMesh)->mesh)

CPolyMeshReaderOF::CPolyMeshReaderOF(void* pRunTime) {
	m_pRunTime = pRunTime;
	ShimTime* t = static_cast<ShimTime*>(pRunTime);
	ShimMesh* sm = new ShimMesh;
	try {
		sm->mesh = lfm::readPolyMesh(t->caseDir + "/constant/polyMesh");
	} catch (const std::exception& e) {
		fatal(e.what());
	}
	m_pMesh = sm;
	t->mesh = sm;
	m_nTimeIndex = 0;
	m_nPointProcAddressingList = sm->mesh.pointProcAddressing;
	m_nFaceProcAddressingList = sm->mesh.faceProcAddressing;
	m_nCellProcAddressingList = sm->mesh.cellProcAddressing;
	m_nBoundaryProcAddressingList = sm->mesh.boundaryProcAddressing;
	m_nCellSubmeshList = sm->mesh.cellSubmesh;
	// faces lying on `empty` patches are not "valid" (polyMeshReaderOF.cpp:47-64)
	m_bValidFaceFlagList.assign((size_t)sm->mesh.nFaces(), true);
	for (const lfm::Patch& p : sm->mesh.patches)
		if (p.type == "empty")
			for (int f = p.startFace; f < p.startFace + p.nFaces; f++) m_bValidFaceFlagList[(size_t)f] = false;
	initializeSubmesh();
}

CPolyMeshReaderOF::~CPolyMeshReaderOF() {
	ShimMesh* sm = static_cast<ShimMesh*>(m_pMesh);
	if (sm) {
		for (ShimField* f : sm->fields) delete f;
		ShimTime* t = static_cast<ShimTime*>(m_pRunTime);
		if (t && t->mesh == sm) t->mesh = nullptr;
		delete sm;
	}
}

int CPolyMeshReaderOF::getPointCount() const { return MESH.nPoints(); }
void CPolyMeshReaderOF::getPoint(const int i, double* p) const {
	for (int k = 0; k < 3; k++) p[k] = MESH.points[(size_t)i * 3 + k];
}
int CPolyMeshReaderOF::getPointProcAddressing(const int i) const {
	return (i < (int)m_nPointProcAddressingList.size()) ? m_nPointProcAddressingList[i] : i;
}
int CPolyMeshReaderOF::getFaceCount() const { return MESH.nFaces(); }
void CPolyMeshReaderOF::getFaceAreaNormal(const int f, double* n) const {
	for (int k = 0; k < 3; k++) n[k] = MESH.faceAreas[(size_t)f * 3 + k];
}
void CPolyMeshReaderOF::getFaceCenter(const int f, double* x) const {
	for (int k = 0; k < 3; k++) x[k] = MESH.faceCentres[(size_t)f * 3 + k];
}
int CPolyMeshReaderOF::getFaceOwner(const int f) const { return MESH.owner[(size_t)f]; }
int CPolyMeshReaderOF::getFaceNeighbour(const int f) const {
	if (f < MESH.nInternalFaces()) return MESH.neighbour[(size_t)f];
	const int c = getCyclicFaceIndex(f);
	return (c == f) ? -1 : MESH.owner[(size_t)c];
}
int CPolyMeshReaderOF::getFaceBoundary(const int f) const { return MESH.whichPatch(f); }
int CPolyMeshReaderOF::getFaceIndexInsideBoundary(const int f) const {
	const int b = MESH.whichPatch(f);
	return f - (b == -1 ? 0 : MESH.patches[(size_t)b].startFace);
}
int CPolyMeshReaderOF::getFacePointCount(const int f) const { return MESH.facePointCount(f); }
int CPolyMeshReaderOF::getFaceProcAddressing(const int f) const {
	return (f < (int)m_nFaceProcAddressingList.size()) ? m_nFaceProcAddressingList[f] : f;
}
int CPolyMeshReaderOF::getFacePointIndex(const int f, const int i) const {
	return MESH.facePoints[(size_t)MESH.faceOffsets[(size_t)f] + i];
}
std::vector<int> CPolyMeshReaderOF::getFacePointIndexList(const int f) const {
	return std::vector<int>(MESH.facePoints.begin() + MESH.faceOffsets[(size_t)f], MESH.facePoints.begin() + MESH.faceOffsets[(size_t)f + 1]);
}
int CPolyMeshReaderOF::getCyclicFaceIndex(const int f) const {
	const int b = MESH.whichPatch(f);
	if (b == -1 || !MESH.patches[(size_t)b].coupled()) return f;
	const int nb = MESH.patches[(size_t)b].neighbPatchID;
	if (nb == -1) return f;
	return f - MESH.patches[(size_t)b].startFace + MESH.patches[(size_t)nb].startFace;
}
int CPolyMeshReaderOF::getFaceId(const int f) const {
	const int tag = getBoundaryTag(getFaceBoundary(f));
	if (tag >= -1 && tag <= 1) return std::abs(getFaceProcAddressing(f));
	return getFaceIndexInsideBoundary(f);
}

int CPolyMeshReaderOF::getCellCount() const { return MESH.nCells; }
double CPolyMeshReaderOF::getCellVolume(const int c) const { return MESH.cellVolumes[(size_t)c]; }
void CPolyMeshReaderOF::getCellCenter(const int c, double* x) const {
	for (int k = 0; k < 3; k++) x[k] = MESH.cellCentres[(size_t)c * 3 + k];
}
int CPolyMeshReaderOF::getCellFaceCount(const int c) const { return MESH.cellFaceOffsets[(size_t)c + 1] - MESH.cellFaceOffsets[(size_t)c]; }
int CPolyMeshReaderOF::getCellFaceIndex(const int c, const int i) const { return MESH.cellFaces[(size_t)MESH.cellFaceOffsets[(size_t)c] + i]; }
bool CPolyMeshReaderOF::getCellFaceOwner(const int c, const int i) const { return MESH.owner[(size_t)getCellFaceIndex(c, i)] == c; }
int CPolyMeshReaderOF::getCellProcAddressing(const int c) const {
	return (c < (int)m_nCellProcAddressingList.size()) ? m_nCellProcAddressingList[c] : c;
}
std::vector<int> CPolyMeshReaderOF::getCellPointList(const int c) const {
	std::vector<int> out;
	for (int i = MESH.cellFaceOffsets[(size_t)c]; i < MESH.cellFaceOffsets[(size_t)c + 1]; i++) {
		const int f = MESH.cellFaces[(size_t)i];
		for (int k = MESH.faceOffsets[(size_t)f]; k < MESH.faceOffsets[(size_t)f + 1]; k++) {
			const int p = MESH.facePoints[(size_t)k];
			bool isNew = true;
			for (size_t j = 0; j < out.size() && isNew; j++) isNew = (out[j] != p);
			if (isNew) out.push_back(p);
		}
	}
	return out;
}
int CPolyMeshReaderOF::getCellValidFaceCount(const int c) const {
	int n = 0;
	for (int i = MESH.cellFaceOffsets[(size_t)c]; i < MESH.cellFaceOffsets[(size_t)c + 1]; i++)
		if (m_bValidFaceFlagList[(size_t)MESH.cellFaces[(size_t)i]]) n++;
	return n;
}
int CPolyMeshReaderOF::getCellValidFaceIndex(const int c, const int idx) const {
	int n = 0;
	for (int i = MESH.cellFaceOffsets[(size_t)c]; i < MESH.cellFaceOffsets[(size_t)c + 1]; i++)
		if (m_bValidFaceFlagList[(size_t)MESH.cellFaces[(size_t)i]]) {
			if (n == idx) return MESH.cellFaces[(size_t)i];
			n++;
		}
	return -1;
}
double CPolyMeshReaderOF::getCellScalar(const int c, const int fi) const {
	return static_cast<ShimField*>(m_pVolScalarFieldList[(size_t)fi])->v[(size_t)c];
}
void CPolyMeshReaderOF::getCellVector(const int c, const int fi, double* v) const {
	const ShimField* f = static_cast<ShimField*>(m_pVolVectorFieldList[(size_t)fi]);
	for (int k = 0; k < 3; k++) v[k] = f->v[(size_t)c * 3 + k];
}

int CPolyMeshReaderOF::getBoundaryCount() const { return (int)MESH.patches.size(); }
std::string CPolyMeshReaderOF::getBoundaryName(const int b) const { return MESH.patches[(size_t)b].name; }
std::string CPolyMeshReaderOF::getBoundaryType(const int b) const { return MESH.patches[(size_t)b].type; }
int CPolyMeshReaderOF::getBoundaryFaceStart(const int b) const { return MESH.patches[(size_t)b].startFace; }
int CPolyMeshReaderOF::getBoundaryFaceEnd(const int b) const { return MESH.patches[(size_t)b].startFace + MESH.patches[(size_t)b].nFaces - 1; }
std::vector<int> CPolyMeshReaderOF::getBoundaryFaceCellList(const int b) const {
	const lfm::Patch& p = MESH.patches[(size_t)b];
	return std::vector<int>(MESH.owner.begin() + p.startFace, MESH.owner.begin() + p.startFace + p.nFaces);
}
int CPolyMeshReaderOF::getBoundaryProcAddressing(const int b) const {
	return (b < (int)m_nBoundaryProcAddressingList.size()) ? m_nBoundaryProcAddressingList[b] : b;
}
int CPolyMeshReaderOF::getBoundaryProcessorRank(const int b) const {
	if (b < MESH.nNonProcessor()) return -1;
	return MESH.patches[(size_t)b].neighbProcNo;
}
int CPolyMeshReaderOF::getBoundaryTag(const int b) const {
	return MESH.processorTag(b);   // tag() * (2 owner() - 1); plain processor patches: tag() == UPstream::msgType() == 1
}
int CPolyMeshReaderOF::getBoundaryCyclicPairIndex(const int b) const {
	return MESH.patches[(size_t)b].coupled() ? -1 : MESH.patches[(size_t)b].neighbPatchID;
}
std::vector<int> CPolyMeshReaderOF::getNeighbourCellList(const int b) const {
	const lfm::Patch& p = MESH.patches[(size_t)b];
	std::vector<int> out;
	if (p.isCyclic() && p.neighbPatchID >= 0) {
		const lfm::Patch& q = MESH.patches[(size_t)p.neighbPatchID];
		out.assign(MESH.owner.begin() + q.startFace, MESH.owner.begin() + q.startFace + q.nFaces);
	}
	return out;
}

void CPolyMeshReaderOF::initializeSubmesh() {
	const lfm::PolyMesh& mesh = MESH;
	const unsigned nPointCount = (unsigned)mesh.nPoints();
	const unsigned nFaceCount = (unsigned)mesh.nFaces();
	const unsigned nCellCount = (unsigned)mesh.nCells;
	if (m_nCellSubmeshList.size() != nCellCount) {
		std::vector<bool> interiorPoint(nPointCount, true);
		m_nCellSubmeshList.assign(nCellCount, 1);
		for (const lfm::Patch& p : mesh.patches) {
			if (p.type == "empty") continue;
			for (int f = p.startFace; f < p.startFace + p.nFaces; f++)
				for (int k = mesh.faceOffsets[(size_t)f]; k < mesh.faceOffsets[(size_t)f + 1]; k++) interiorPoint[(size_t)mesh.facePoints[(size_t)k]] = false;
		}
		for (unsigned c = 0; c < nCellCount; c++) {
			bool interior = true;
			for (int i = mesh.cellFaceOffsets[c]; i < mesh.cellFaceOffsets[c + 1] && interior; i++) {
				const int f = mesh.cellFaces[(size_t)i];
				for (int k = mesh.faceOffsets[(size_t)f]; k < mesh.faceOffsets[(size_t)f + 1] && interior; k++)
					interior = interiorPoint[(size_t)mesh.facePoints[(size_t)k]];
			}
			m_nCellSubmeshList[c] = interior ? 1 : 0;
		}
	}
	m_nPointMaskList.assign(nPointCount, 0);
	m_nFaceMaskList.assign(nFaceCount, 0);
	m_nCellMaskList.assign(nCellCount, 0);
	m_nSubmeshCount = 0;
	for (unsigned c = 0; c < nCellCount; c++) {
		const int sm = m_nCellSubmeshList[c];
		if (sm >= m_nSubmeshCount) m_nSubmeshCount = sm + 1;
		const int mask = getMask(sm);
		m_nCellMaskList[c] |= mask;
		for (int i = mesh.cellFaceOffsets[c]; i < mesh.cellFaceOffsets[c + 1]; i++) {
			const int f = mesh.cellFaces[(size_t)i];
			m_nFaceMaskList[(size_t)f] |= mask;
			for (int k = mesh.faceOffsets[(size_t)f]; k < mesh.faceOffsets[(size_t)f + 1]; k++) m_nPointMaskList[(size_t)mesh.facePoints[(size_t)k]] |= mask;
		}
	}
	m_nSubmeshPointMappingList.assign((size_t)m_nSubmeshCount, {});
	m_nSubmeshFaceMappingList.assign((size_t)m_nSubmeshCount, {});
	m_nSubmeshCellMappingList.assign((size_t)m_nSubmeshCount, {});
	for (int s = 0; s < m_nSubmeshCount; s++) {
		const std::vector<int> pl = getSubmeshPointIndexList(s);
		for (unsigned i = 0; i < pl.size(); i++) m_nSubmeshPointMappingList[(size_t)s].insert({pl[i], (int)i});
		const std::vector<int> fl = getSubmeshFaceIndexList(s);
		for (unsigned i = 0; i < fl.size(); i++) m_nSubmeshFaceMappingList[(size_t)s].insert({fl[i], (int)i});
		const std::vector<int> cl = getSubmeshCellIndexList(s);
		for (unsigned i = 0; i < cl.size(); i++) m_nSubmeshCellMappingList[(size_t)s].insert({cl[i], (int)i});
	}
}

std::vector<double> CPolyMeshReaderOF::getCellDistanceFromBoundary(const int, const std::vector<int>) const {
	fatal("getCellDistanceFromBoundary is not provided by the shim (not on the solve path)");
}

double CPolyMeshReaderOF::getLocalityScore(const int nMethod) const {
	const lfm::PolyMesh& mesh = MESH;
	double score = 0;
	for (int c = 0; c < mesh.nCells; c++) {
		double cs = 0;
		int nn = 0;
		for (int i = mesh.cellFaceOffsets[(size_t)c]; i < mesh.cellFaceOffsets[(size_t)c + 1]; i++) {
			const int f = mesh.cellFaces[(size_t)i];
			if (f < mesh.nInternalFaces()) {
				const double d = 1.0 * std::abs(mesh.neighbour[(size_t)f] - c) / mesh.nCells;
				cs += (nMethod == 0) ? d : d * d;
				nn++;
			}
		}
		score += cs / nn;
	}
	return score / mesh.nCells;
}

static int readField(ShimMesh* sm, ShimTime* t, const std::string& name, int nComp, std::vector<void*>& list,
                     std::vector<std::string>& names) {
	ShimField* f = new ShimField;
	f->name = name;
	f->nComp = nComp;
	try {
		f->v = lfm::readVolField(t->caseDir + "/" + t->timeName() + "/" + name, sm->mesh.nCells, nComp);
	} catch (const std::exception& e) {
		fatal(e.what());
	}
	sm->fields.push_back(f);
	names.push_back(name);
	list.push_back(f);
	return (int)list.size() - 1;
}

int CPolyMeshReaderOF::readScalarField(const std::string sFieldName) {
	int idx = getScalarFieldIndex(sFieldName);
	if (idx != -1) return idx;
	return readField(static_cast<ShimMesh*>(m_pMesh), static_cast<ShimTime*>(m_pRunTime), sFieldName, 1, m_pVolScalarFieldList, m_sScalarFieldList);
}
int CPolyMeshReaderOF::readVectorField(const std::string sFieldName) {
	int idx = getVectorFieldIndex(sFieldName);
	if (idx != -1) return idx;
	return readField(static_cast<ShimMesh*>(m_pMesh), static_cast<ShimTime*>(m_pRunTime), sFieldName, 3, m_pVolVectorFieldList, m_sVectorFieldList);
}
int CPolyMeshReaderOF::createScalarField(const std::string sFieldName) {
	int idx = getScalarFieldIndex(sFieldName);
	if (idx != -1) return idx;
	ShimMesh* sm = static_cast<ShimMesh*>(m_pMesh);
	ShimField* f = new ShimField;
	f->name = sFieldName;
	f->nComp = 1;
	f->v.assign((size_t)sm->mesh.nCells, 0.0);
	sm->fields.push_back(f);
	m_sScalarFieldList.push_back(sFieldName);
	m_pVolScalarFieldList.push_back(f);
	return (int)m_pVolScalarFieldList.size() - 1;
}

#define SF(i) (static_cast<ShimField*>(m_pVolScalarFieldList[(size_t)(i)])->v)
#define VF(i) (static_cast<ShimField*>(m_pVolVectorFieldList[(size_t)(i)])->v)

void CPolyMeshReaderOF::updateScalarField(const int fi, const std::vector<double>& s, const int sm, const double fac) {
	std::vector<double>& f = SF(fi);
	if (sm == -1) {
		for (size_t c = 0; c < f.size(); c++) f[c] = s[c] * fac;
	} else {
		const std::vector<int> cl = getSubmeshCellIndexList(sm);
		for (size_t i = 0; i < cl.size(); i++) f[(size_t)cl[i]] = s[i] * fac;
	}
}
void CPolyMeshReaderOF::updateScalarField(const int fi, const std::vector<float>& s, const int sm, const float fac) {
	std::vector<double>& f = SF(fi);
	if (sm == -1) {
		for (size_t c = 0; c < f.size(); c++) f[c] = s[c] * fac;
	} else {
		const std::vector<int> cl = getSubmeshCellIndexList(sm);
		for (size_t i = 0; i < cl.size(); i++) f[(size_t)cl[i]] = s[i] * fac;
	}
}
void CPolyMeshReaderOF::updateScalarFieldSqrt(const int fi, const std::vector<double>& s, const int sm, const double fac) {
	std::vector<double>& f = SF(fi);
	if (sm == -1) {
		for (size_t c = 0; c < f.size(); c++) f[c] = std::sqrt(s[c] * fac);
	} else {
		const std::vector<int> cl = getSubmeshCellIndexList(sm);
		for (size_t i = 0; i < cl.size(); i++) f[(size_t)cl[i]] = std::sqrt(s[i] * fac);
	}
}
void CPolyMeshReaderOF::updateScalarFieldSqrt(const int fi, const std::vector<float>& s, const int sm, const float fac) {
	std::vector<double>& f = SF(fi);
	if (sm == -1) {
		for (size_t c = 0; c < f.size(); c++) f[c] = std::sqrt(s[c] * fac);
	} else {
		const std::vector<int> cl = getSubmeshCellIndexList(sm);
		for (size_t i = 0; i < cl.size(); i++) f[(size_t)cl[i]] = std::sqrt(s[i] * fac);
	}
}
void CPolyMeshReaderOF::updateScalarField(const int fi, const double v, const int c) { SF(fi)[(size_t)c] = v; }
void CPolyMeshReaderOF::updateScalarField(const int fi, const float v, const int c) { SF(fi)[(size_t)c] = v; }
void CPolyMeshReaderOF::updateVectorField(const int fi, const std::vector<double>& s, const int sm) {
	std::vector<double>& f = VF(fi);
	if (sm == -1) {
		for (size_t i = 0; i < f.size(); i++) f[i] = s[i];
	} else {
		const std::vector<int> cl = getSubmeshCellIndexList(sm);
		for (size_t i = 0; i < cl.size(); i++)
			for (int k = 0; k < 3; k++) f[(size_t)cl[i] * 3 + k] = s[i * 3 + k];
	}
}
void CPolyMeshReaderOF::updateVectorField(const int fi, const std::vector<float>& s, const int sm) {
	std::vector<double>& f = VF(fi);
	if (sm == -1) {
		for (size_t i = 0; i < f.size(); i++) f[i] = s[i];
	} else {
		const std::vector<int> cl = getSubmeshCellIndexList(sm);
		for (size_t i = 0; i < cl.size(); i++)
			for (int k = 0; k < 3; k++) f[(size_t)cl[i] * 3 + k] = s[i * 3 + k];
	}
}
void CPolyMeshReaderOF::updateVectorField(const int fi, const double* v, const int c) {
	for (int k = 0; k < 3; k++) VF(fi)[(size_t)c * 3 + k] = v[k];
}
void CPolyMeshReaderOF::updateVectorField(const int fi, const float* v, const int c) {
	for (int k = 0; k < 3; k++) VF(fi)[(size_t)c * 3 + k] = v[k];
}
int CPolyMeshReaderOF::getScalarFieldIndex(const std::string& n) const {
	for (unsigned i = 0; i < m_sScalarFieldList.size(); i++)
		if (n.compare(m_sScalarFieldList[i]) == 0) return (int)i;
	return -1;
}
int CPolyMeshReaderOF::getVectorFieldIndex(const std::string& n) const {
	for (unsigned i = 0; i < m_sVectorFieldList.size(); i++)
		if (n.compare(m_sVectorFieldList[i]) == 0) return (int)i;
	return -1;
}
