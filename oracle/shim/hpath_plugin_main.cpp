// Driver of the reference's hpathRenumber plugin (/root/reference/hpathRenumber/hpathRenumber.C, compiled UNCHANGED against
// oracle/shim/openfoam_stub/): reads a case's polyMesh with the OpenFOAM-free reader, hands the plugin a Foam::polyMesh
// stand-in built from it and writes the order it returns (int32, order[new] = old).  TEST INFRASTRUCTURE ONLY
// (tests/test_hpath_vs_plugin.py pins lfm_public_b200/host/hpath.cpp against it).
//
//     hpath_plugin <case>/constant/polyMesh <order.bin>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "foam_io.h"
#include "hpathRenumber.H"

int main(int argc, char** argv) {
	if (argc < 3) {
		fprintf(stderr, "usage: hpath_plugin <polyMesh dir> <order.bin>\n");
		return 2;
	}
	try {
		lfm::PolyMesh m = lfm::readPolyMesh(argv[1]);
		m.finalize();
		Foam::polyMesh fm;
		for (int p = 0; p < m.nPoints(); p++) fm.points_.push_back(Foam::point(m.points[3 * p], m.points[3 * p + 1], m.points[3 * p + 2]));
		for (int f = 0; f < m.nFaces(); f++) {
			Foam::face fc;
			for (int k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; k++) fc.push_back(m.facePoints[k]);
			fm.faces_.push_back(fc);
			fm.owner_.push_back(m.owner[f]);
		}
		for (int f = 0; f < m.nInternalFaces(); f++) fm.neighbour_.push_back(m.neighbour[f]);
		fm.calcCells(m.nCells);
		// the reader's own cells() must be what OpenFOAM would build (the restatement walks it)
		for (int c = 0; c < m.nCells; c++) {
			const int n = m.cellFaceOffsets[c + 1] - m.cellFaceOffsets[c];
			if (n != fm.cells_[c].size()) throw std::runtime_error("cells(): face count differs from primitiveMesh::calcCells");
			for (int k = 0; k < n; k++)
				if (m.cellFaces[m.cellFaceOffsets[c] + k] != fm.cells_[c][k]) throw std::runtime_error("cells(): face order differs from primitiveMesh::calcCells");
		}
		for (int c = 0; c < m.nCells; c++) fm.cellCentres_.push_back(Foam::point(m.cellCentres[3 * c], m.cellCentres[3 * c + 1], m.cellCentres[3 * c + 2]));
		for (const lfm::Patch& p : m.patches) fm.boundary_.append(Foam::polyPatch(p.type, p.startFace, p.nFaces));
		Foam::dictionary dict;
		Foam::hpathRenumber plugin(dict);
		const Foam::labelList order = plugin.renumber(fm, fm.cellCentres());
		FILE* out = fopen(argv[2], "wb");
		if (!out) throw std::runtime_error("cannot write the order file");
		for (int c = 0; c < order.size(); c++) {
			const int32_t v = order[c];
			fwrite(&v, sizeof v, 1, out);
		}
		fclose(out);
		return 0;
	} catch (const std::exception& e) {
		fprintf(stderr, "hpath_plugin: %s\n", e.what());
		return 1;
	}
}
