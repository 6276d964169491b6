// C ABI over the host-side case setup (include/lfmhost.h).
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include "flatten.h"
#include "lfmhost.h"

namespace {
thread_local std::string g_err;
int setErr(const std::exception& e) {
	g_err = e.what();
	return 1;
}
}  // namespace

struct lfmhost_case {
	lfm::PolyMesh mesh;
	lfm::CaseOptions opts;
	lfm::Fields fields;
	std::unique_ptr<lfm::FlatMesh> flat;
	std::vector<std::vector<char>> exports;
	int rank = -1, nRanks = 1;
};

static void toC(const lfm::CaseOptions& o, lfmhost_opts* c) {
	c->comm_type = o.commType;
	c->halo_comm_type = o.haloCommType;
	c->double_precision = o.doublePrecision;
	c->have_average = o.haveAverage;
	c->have_forces = o.haveForces;
	c->have_residual = o.haveResidual;
	c->save_forces_step = o.saveForcesStep;
	c->print_info_freq = o.printInfoFreq;
	c->t_start_average = o.tStartAverage;
	c->cfl_max = o.cflMax;
	c->solver = o.solver;
	c->dimension = o.dimension;
	c->rk_order = o.rkOrder;
	c->minmod = o.minmod;
	c->p_inf = o.pInf;
	c->T_inf = o.TInf;
	for (int k = 0; k < 3; k++) c->U_inf[k] = o.UInf[k];
	c->Ls = o.Ls;
	c->mach = o.Mach;
	c->K = o.K;
	c->Cp = o.Cp;
	c->mol_weight = o.molWeight;
	c->mu0 = o.mu0;
	c->Pr = o.Pr;
	c->laminar = o.laminar;
	c->start_time = o.startTime;
	c->end_time = o.endTime;
	c->delta_t = o.deltaT;
	c->write_interval = o.writeInterval;
	c->adjust_time_step = o.adjustTimeStep;
	c->time_precision = o.timePrecision;
	c->write_precision = o.writePrecision;
}

static void fromC(const lfmhost_opts* c, lfm::CaseOptions& o) {
	o.commType = c->comm_type;
	o.haloCommType = c->halo_comm_type;
	o.doublePrecision = c->double_precision != 0;
	o.haveAverage = c->have_average != 0;
	o.haveForces = c->have_forces != 0;
	o.haveResidual = c->have_residual != 0;
	o.saveForcesStep = c->save_forces_step;
	o.printInfoFreq = c->print_info_freq;
	o.tStartAverage = c->t_start_average;
	o.cflMax = c->cfl_max;
	o.solver = c->solver;
	o.dimension = c->dimension;
	o.rkOrder = c->rk_order;
	o.minmod = c->minmod != 0;
	o.pInf = c->p_inf;
	o.TInf = c->T_inf;
	for (int k = 0; k < 3; k++) o.UInf[k] = c->U_inf[k];
	o.Ls = c->Ls;
	o.Mach = c->mach;
	o.K = c->K;
	o.Cp = c->Cp;
	o.molWeight = c->mol_weight;
	o.mu0 = c->mu0;
	o.Pr = c->Pr;
	o.laminar = c->laminar != 0;
	o.startTime = c->start_time;
	o.endTime = c->end_time;
	o.deltaT = c->delta_t;
	o.writeInterval = c->write_interval;
	o.adjustTimeStep = c->adjust_time_step != 0;
	o.timePrecision = c->time_precision;
	o.writePrecision = c->write_precision;
}

extern "C" {

const char* lfmhost_last_error(void) { return g_err.c_str(); }

void lfmhost_default_opts(lfmhost_opts* o) {
	lfm::CaseOptions d;
	toC(d, o);
}

int lfmhost_open_case(const char* case_dir, int rank, int n_ranks, lfmhost_case** out) {
	try {
		std::unique_ptr<lfmhost_case> c(new lfmhost_case);
		c->rank = rank < 0 ? 0 : rank;
		c->nRanks = n_ranks < 1 ? 1 : n_ranks;
		const std::string root(case_dir);
		const std::string dir = rank < 0 ? root : root + "/processor" + std::to_string(rank);
		c->opts = lfm::readCaseOptions(root);
		c->mesh = lfm::readPolyMesh(dir + "/constant/polyMesh");
		c->fields = lfm::readFields(dir, c->opts, c->mesh.nCells);
		c->flat.reset(new lfm::FlatMesh(c->mesh, c->fields, c->opts, c->rank, c->nRanks));
		c->exports.resize((size_t)c->flat->neighbourCount());
		*out = c.release();
		return 0;
	} catch (const std::exception& e) {
		return setErr(e);
	}
}

int lfmhost_case_from_arrays(const lfmhost_mesh_in* in, const lfmhost_opts* opts, const double* p, const double* T, const double* U,
                             const double* alpha, int rank, int n_ranks, lfmhost_case** out) {
	try {
		std::unique_ptr<lfmhost_case> c(new lfmhost_case);
		c->rank = rank < 0 ? 0 : rank;
		c->nRanks = n_ranks < 1 ? 1 : n_ranks;
		fromC(opts, c->opts);
		lfm::PolyMesh& m = c->mesh;
		m.points.assign(in->points, in->points + (size_t)in->n_points * 3);
		m.faceOffsets.resize((size_t)in->n_faces + 1);
		m.facePoints.reserve((size_t)in->n_faces * 4);
		m.faceOffsets[0] = 0;
		for (int f = 0; f < in->n_faces; f++) {
			for (int k = 0; k < 4; k++) {
				const int v = in->faces[(size_t)f * 4 + k];
				if (v >= 0) m.facePoints.push_back(v);
			}
			m.faceOffsets[(size_t)f + 1] = (int)m.facePoints.size();
		}
		m.owner.assign(in->owner, in->owner + in->n_faces);
		m.neighbour.assign(in->neighbour, in->neighbour + in->n_internal);
		m.nCells = in->n_cells;
		for (int b = 0; b < in->n_patches; b++) {
			lfm::Patch pt;
			pt.name = in->patch_name[b];
			pt.type = in->patch_type[b];
			pt.nFaces = in->patch_nfaces[b];
			pt.startFace = in->patch_start[b];
			pt.neighbourPatch = in->patch_nbr_name ? in->patch_nbr_name[b] : "";
			pt.myProcNo = in->patch_my_proc ? in->patch_my_proc[b] : -1;
			pt.neighbProcNo = in->patch_nbr_proc ? in->patch_nbr_proc[b] : -1;
			pt.referPatch = in->patch_refer_name ? in->patch_refer_name[b] : "";
			m.patches.push_back(pt);
		}
		if (in->face_proc_addressing) m.faceProcAddressing.assign(in->face_proc_addressing, in->face_proc_addressing + in->n_faces);
		if (in->cell_submesh) m.cellSubmesh.assign(in->cell_submesh, in->cell_submesh + in->n_cells);
		m.finalize();
		const size_t nc = (size_t)m.nCells;
		c->fields.p.assign(p, p + nc);
		c->fields.T.assign(T, T + nc);
		c->fields.U.assign(U, U + nc * 3);
		if (alpha) c->fields.alpha.assign(alpha, alpha + nc);
		c->flat.reset(new lfm::FlatMesh(c->mesh, c->fields, c->opts, c->rank, c->nRanks));
		c->exports.resize((size_t)c->flat->neighbourCount());
		*out = c.release();
		return 0;
	} catch (const std::exception& e) {
		return setErr(e);
	}
}

int lfmhost_get_opts(const lfmhost_case* c, lfmhost_opts* out) {
	toC(c->opts, out);
	return 0;
}
int lfmhost_nbr_count(const lfmhost_case* c) { return c->flat->neighbourCount(); }
int lfmhost_nbr_rank(const lfmhost_case* c, int i) { return c->flat->neighbourRank(i); }

int lfmhost_export(lfmhost_case* c, int i, const void** data, size_t* bytes) {
	try {
		c->exports[(size_t)i] = c->flat->exportFor(i);
		*data = c->exports[(size_t)i].data();
		*bytes = c->exports[(size_t)i].size();
		return 0;
	} catch (const std::exception& e) {
		return setErr(e);
	}
}
int lfmhost_import(lfmhost_case* c, int i, const void* data, size_t bytes) {
	try {
		c->flat->importFrom(i, (const char*)data, bytes);
		return 0;
	} catch (const std::exception& e) {
		return setErr(e);
	}
}
int lfmhost_finish(lfmhost_case* c) {
	try {
		c->flat->finish();
		return 0;
	} catch (const std::exception& e) {
		return setErr(e);
	}
}
const lfmgpu_desc* lfmhost_desc(const lfmhost_case* c) { return &c->flat->desc(); }

int lfmhost_geometry(const lfmhost_case* c, double* fa, double* fc, double* cc, double* cv) {
	const lfm::PolyMesh& m = c->mesh;
	if (fa) memcpy(fa, m.faceAreas.data(), m.faceAreas.size() * sizeof(double));
	if (fc) memcpy(fc, m.faceCentres.data(), m.faceCentres.size() * sizeof(double));
	if (cc) memcpy(cc, m.cellCentres.data(), m.cellCentres.size() * sizeof(double));
	if (cv) memcpy(cv, m.cellVolumes.data(), m.cellVolumes.size() * sizeof(double));
	return 0;
}
int lfmhost_hpath_order(const lfmhost_case* c, int32_t* order, double* stats) {
	try {
		double st[4] = {0, 0, 0, 0};
		const std::vector<int> o = lfm::hpathOrder(c->mesh, st);
		if ((int)o.size() != c->mesh.nCells) throw std::runtime_error("hpath: the order does not cover the mesh");
		for (size_t i = 0; i < o.size(); i++) order[i] = o[i];
		if (stats)
			for (int k = 0; k < 4; k++) stats[k] = st[k];
		return 0;
	} catch (const std::exception& e) {
		return setErr(e);
	}
}
int lfmhost_mesh_sizes(const lfmhost_case* c, int32_t* np, int32_t* nf, int32_t* ni, int32_t* nc) {
	if (np) *np = c->mesh.nPoints();
	if (nf) *nf = c->mesh.nFaces();
	if (ni) *ni = c->mesh.nInternalFaces();
	if (nc) *nc = c->mesh.nCells;
	return 0;
}

int lfmhost_write_field(const lfmhost_case* c, const char* path, const char* name, const double* values, int n_comp, int precision) {
	try {
		const std::vector<int>& gid = c->flat->cellGid();
		std::vector<double> v((size_t)c->mesh.nCells * n_comp);
		for (size_t t = 0; t < gid.size(); t++)
			for (int k = 0; k < n_comp; k++) v[(size_t)gid[t] * n_comp + k] = values[t * n_comp + k];
		lfm::writeVolField(path, name, c->mesh, v, n_comp, precision);
		return 0;
	} catch (const std::exception& e) {
		return setErr(e);
	}
}

int lfmhost_close(lfmhost_case* c) {
	delete c;
	return 0;
}

}  // extern "C"
