// mini-MPI implementation (see mpi.h).  TEST INFRASTRUCTURE ONLY: lets the unmodified reference run
// multi-rank on one host without an MPI installation.
//
//  * LFM_MPI_NP=<n>        number of ranks; rank 0 is the launching process, ranks 1..n-1 are forked in MPI_Init
//  * LFM_MPI_DUMP_DIR=<d>  every float/double point-to-point payload is appended to <d>/send_r<src>_to<dst>.bin as
//                          [int32 tag][int32 nbytes][payload]; used to pin the halo wire format bit-exactly
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <sched.h>
#include <signal.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <deque>
#include <string>
#include <vector>

namespace {

struct Shared {
	std::atomic<int> barrier_count;
	std::atomic<int> barrier_sense;
	std::atomic<int> abort_code;
	char pad[64];
	// per-rank collective slots follow
};

const size_t SLOT_BYTES = 1 << 16;

int g_rank = 0, g_size = 1;
bool g_init = false, g_final = false;
Shared* g_sh = nullptr;
char* g_slots = nullptr;
int g_local_sense = 0;
std::vector<int> g_fd;       // fd to each peer (-1 for self)
std::vector<pid_t> g_children;
std::string g_dump_dir;
std::vector<FILE*> g_dump;   // per destination

struct TypeInfo {
	size_t extent;
	bool is_fp;
};
std::vector<TypeInfo> g_types = {{0, false}, {1, false}, {4, false}, {4, false}, {2, false}, {4, true}, {8, true}, {4, false}, {1, false}};

enum Kind { SEND, RECV };
struct Req {
	Kind kind;
	int peer, tag;
	char* buf;
	size_t nbytes;
	bool persistent = false, active = false, done = true, in_use = false;
};
std::vector<Req> g_reqs;

struct Hdr {
	int32_t tag, nbytes;
};
struct SendState {
	int req;
	Hdr hdr;
	size_t off;   // bytes written of hdr+payload
};
struct Unexpected {
	Hdr hdr;
	std::vector<char> data;
};
struct Peer {
	std::deque<SendState> sendq;
	std::deque<int> recvq;            // posted receives (FIFO)
	std::deque<Unexpected> unexpected;
	// incoming message being assembled
	bool have_hdr = false;
	Hdr hdr;
	size_t hdr_off = 0, body_off = 0;
	int cur_req = -1;                 // posted receive the body is streamed into, -1: into `stash`
	std::vector<char> stash;
	bool body_active = false;
};
std::vector<Peer> g_peers;

// ---- one-sided communication, general active target synchronisation (post / start / complete / wait) ----------------------
// The ranks are separate processes without a common mapping of the window memory, so an exposure epoch is emulated with
// messages on the same sockets: MPI_Win_post sends the window's contents to every origin of the group (the target may not
// change an exposed window before MPI_Win_wait returns, so the copy taken at the post IS what a get of that epoch reads),
// MPI_Get copies out of the copy received from its target, MPI_Win_complete tells every target that the origin is done and
// MPI_Win_wait collects those notes.  The messages carry reserved tags and never enter the point-to-point matching: they
// are queued per (peer, window) as they arrive, whatever the receiver is doing.
const int32_t RMA_SNAP = 0x7f000000, RMA_DONE = 0x7e000000, RMA_KIND = 0x7f000000, RMA_WIN = 0x00ffffff;
inline bool is_rma(int32_t tag) { return (tag & RMA_KIND) == RMA_SNAP || (tag & RMA_KIND) == RMA_DONE; }
struct Window {
	char* base = nullptr;
	size_t size = 0;
	int disp_unit = 1;
	std::vector<int> exposed_to, accessing;                 // groups of the open exposure / access epoch
	std::vector<std::deque<std::vector<char>>> snap;        // [peer]: window copies received, oldest first
	std::vector<int> done;                                  // [peer]: completion notes received and not yet waited for
};
std::vector<Window> g_wins;
std::vector<std::vector<int>> g_groups = {{}};             // handle 0: the world group (filled in MPI_Comm_group)

void rma_deliver(int peer, const Hdr& h, std::vector<char>& data) {
	const size_t w = (size_t)(h.tag & RMA_WIN);
	if (w >= g_wins.size()) {   // a peer may post before this rank has returned from its own MPI_Win_create
		g_wins.resize(w + 1);
	}
	Window& W = g_wins[w];
	if (W.snap.empty()) {
		W.snap.resize((size_t)g_size);
		W.done.assign((size_t)g_size, 0);
	}
	if ((h.tag & RMA_KIND) == RMA_SNAP) {
		W.snap[(size_t)peer].emplace_back();
		W.snap[(size_t)peer].back().swap(data);
	} else {
		W.done[(size_t)peer]++;
	}
}

[[noreturn]] void die(const char* msg) {
	fprintf(stderr, "[mini-mpi rank %d] %s\n", g_rank, msg);
	if (g_sh) g_sh->abort_code.store(99);
	_exit(99);
}

void check_abort() {
	if (g_sh && g_sh->abort_code.load() != 0) _exit(g_sh->abort_code.load());
}

int new_req() {
	for (size_t i = 0; i < g_reqs.size(); i++)
		if (!g_reqs[i].in_use) {
			g_reqs[i] = Req();
			g_reqs[i].in_use = true;
			return (int)i;
		}
	g_reqs.push_back(Req());
	g_reqs.back().in_use = true;
	return (int)g_reqs.size() - 1;
}

void complete_recv(int r) { g_reqs[r].done = true; }

// returns true if any progress was made
bool progress_peer(int p) {
	Peer& P = g_peers[p];
	int fd = g_fd[p];
	bool any = false;
	// sends
	while (!P.sendq.empty()) {
		SendState& s = P.sendq.front();
		Req& rq = g_reqs[s.req];
		const size_t total = sizeof(Hdr) + rq.nbytes;
		bool blocked = false;
		while (s.off < total) {
			const char* src;
			size_t n;
			if (s.off < sizeof(Hdr)) {
				src = (const char*)&s.hdr + s.off;
				n = sizeof(Hdr) - s.off;
			} else {
				src = rq.buf + (s.off - sizeof(Hdr));
				n = total - s.off;
			}
			ssize_t w = write(fd, src, n);
			if (w < 0) {
				if (errno == EAGAIN || errno == EWOULDBLOCK) {
					blocked = true;
					break;
				}
				if (errno == EINTR) continue;
				die("write failed");
			}
			s.off += (size_t)w;
			any = true;
		}
		if (blocked) break;
		rq.done = true;
		P.sendq.pop_front();
	}
	// receives
	while (true) {
		if (!P.have_hdr) {
			ssize_t r = read(fd, (char*)&P.hdr + P.hdr_off, sizeof(Hdr) - P.hdr_off);
			if (r < 0) {
				if (errno == EAGAIN || errno == EWOULDBLOCK) break;
				if (errno == EINTR) continue;
				die("read failed");
			}
			if (r == 0) break;   // peer closed
			P.hdr_off += (size_t)r;
			any = true;
			if (P.hdr_off < sizeof(Hdr)) continue;
			P.have_hdr = true;
			P.hdr_off = 0;
			P.body_off = 0;
			if (!is_rma(P.hdr.tag) && !P.recvq.empty() && P.unexpected.empty()) {
				P.cur_req = P.recvq.front();
				P.recvq.pop_front();
				if ((size_t)P.hdr.nbytes != g_reqs[P.cur_req].nbytes) die("message size mismatch (posted recv vs incoming)");
			} else {
				P.cur_req = -1;
				P.stash.assign((size_t)P.hdr.nbytes, 0);
			}
		}
		char* dst = P.cur_req >= 0 ? g_reqs[P.cur_req].buf : P.stash.data();
		bool blocked = false;
		while (P.body_off < (size_t)P.hdr.nbytes) {
			ssize_t r = read(fd, dst + P.body_off, (size_t)P.hdr.nbytes - P.body_off);
			if (r < 0) {
				if (errno == EAGAIN || errno == EWOULDBLOCK) {
					blocked = true;
					break;
				}
				if (errno == EINTR) continue;
				die("read failed");
			}
			if (r == 0) {
				blocked = true;
				break;
			}
			P.body_off += (size_t)r;
			any = true;
		}
		if (blocked) break;
		if (P.cur_req >= 0) {
			complete_recv(P.cur_req);
		} else if (is_rma(P.hdr.tag)) {
			rma_deliver(p, P.hdr, P.stash);
			P.stash.clear();
		} else {
			Unexpected u;
			u.hdr = P.hdr;
			u.data.swap(P.stash);
			P.unexpected.push_back(std::move(u));
			// a receive may have been posted while this message was being assembled
			while (!P.unexpected.empty() && !P.recvq.empty()) {
				Unexpected& f = P.unexpected.front();
				Req& rq = g_reqs[(size_t)P.recvq.front()];
				if ((size_t)f.hdr.nbytes != rq.nbytes) die("message size mismatch (late match)");
				memcpy(rq.buf, f.data.data(), rq.nbytes);
				rq.done = true;
				P.recvq.pop_front();
				P.unexpected.pop_front();
			}
		}
		P.have_hdr = false;
	}
	return any;
}

bool progress_all() {
	bool any = false;
	for (int p = 0; p < g_size; p++)
		if (p != g_rank) any |= progress_peer(p);
	return any;
}

void idle_wait() {
	check_abort();
	std::vector<pollfd> pf;
	for (int p = 0; p < g_size; p++) {
		if (p == g_rank) continue;
		pollfd x;
		x.fd = g_fd[p];
		x.events = POLLIN;
		if (!g_peers[p].sendq.empty()) x.events |= POLLOUT;
		x.revents = 0;
		pf.push_back(x);
	}
	if (!pf.empty()) poll(pf.data(), pf.size(), 1);
}

void post(int r) {
	Req& rq = g_reqs[r];
	rq.done = false;
	rq.active = true;
	if (rq.peer == g_rank) die("self send/recv is not supported");
	Peer& P = g_peers[rq.peer];
	if (rq.kind == SEND) {
		SendState s;
		s.req = r;
		s.hdr.tag = rq.tag;
		s.hdr.nbytes = (int32_t)rq.nbytes;
		s.off = 0;
		P.sendq.push_back(s);
	} else {
		if (!P.unexpected.empty()) {
			Unexpected& u = P.unexpected.front();
			if ((size_t)u.hdr.nbytes != rq.nbytes) die("message size mismatch (unexpected queue)");
			memcpy(rq.buf, u.data.data(), rq.nbytes);
			P.unexpected.pop_front();
			rq.done = true;
		} else {
			P.recvq.push_back(r);
		}
	}
	progress_peer(rq.peer);
}

void dump_payload(int dest, int tag, const void* buf, size_t nbytes) {
	if (g_dump_dir.empty()) return;
	if (g_dump.empty()) g_dump.assign((size_t)g_size, nullptr);
	if (!g_dump[(size_t)dest]) {
		char path[1024];
		snprintf(path, sizeof path, "%s/send_r%d_to%d.bin", g_dump_dir.c_str(), g_rank, dest);
		g_dump[(size_t)dest] = fopen(path, "wb");
		if (!g_dump[(size_t)dest]) return;
	}
	int32_t h[2] = {(int32_t)tag, (int32_t)nbytes};
	fwrite(h, sizeof h, 1, g_dump[(size_t)dest]);
	fwrite(buf, 1, nbytes, g_dump[(size_t)dest]);
}

void wait_req(int r) {
	while (!g_reqs[r].done) {
		if (!progress_all()) idle_wait();
	}
}

void finish_req(MPI_Request* h) {
	if (*h == MPI_REQUEST_NULL) return;
	Req& rq = g_reqs[*h];
	if (rq.persistent) {
		rq.active = false;
	} else {
		rq.in_use = false;
		*h = MPI_REQUEST_NULL;
	}
}

void barrier_impl() {
	if (g_size == 1) return;
	g_local_sense = !g_local_sense;
	if (g_sh->barrier_count.fetch_add(1) == g_size - 1) {
		g_sh->barrier_count.store(0);
		g_sh->barrier_sense.store(g_local_sense);
	} else {
		int spins = 0;
		while (g_sh->barrier_sense.load() != g_local_sense) {
			progress_all();
			check_abort();
			if (++spins > 64) sched_yield();
		}
	}
}

template <typename T>
void reduce_t(T* acc, const T* x, int n, MPI_Op op) {
	for (int i = 0; i < n; i++) {
		if (op == MPI_SUM) acc[i] += x[i];
		else if (op == MPI_MAX) acc[i] = x[i] > acc[i] ? x[i] : acc[i];
		else acc[i] = x[i] < acc[i] ? x[i] : acc[i];
	}
}

void reduce_any(void* acc, const void* x, int n, MPI_Datatype t, MPI_Op op) {
	switch (t) {
		case MPI_INT: reduce_t((int*)acc, (const int*)x, n, op); break;
		case MPI_UNSIGNED: reduce_t((unsigned*)acc, (const unsigned*)x, n, op); break;
		case MPI_FLOAT: reduce_t((float*)acc, (const float*)x, n, op); break;
		case MPI_DOUBLE: reduce_t((double*)acc, (const double*)x, n, op); break;
		default: die("unsupported datatype in reduction");
	}
}

}  // namespace

extern "C" {

int MPI_Init(int*, char***) {
	if (g_init) return MPI_SUCCESS;
	g_init = true;
	const char* np = getenv("LFM_MPI_NP");
	g_size = np ? atoi(np) : 1;
	if (g_size < 1) g_size = 1;
	const char* dd = getenv("LFM_MPI_DUMP_DIR");
	if (dd) g_dump_dir = dd;
	g_rank = 0;
	g_fd.assign((size_t)g_size, -1);
	g_peers.assign((size_t)g_size, Peer());
	if (g_size == 1) return MPI_SUCCESS;
	size_t bytes = sizeof(Shared) + (size_t)g_size * SLOT_BYTES;
	void* m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
	if (m == MAP_FAILED) die("mmap failed");
	g_sh = new (m) Shared();
	g_sh->barrier_count.store(0);
	g_sh->barrier_sense.store(0);
	g_sh->abort_code.store(0);
	g_slots = (char*)m + sizeof(Shared);
	// socket pair per unordered pair (i<j): sv[i][j] used by i, sv[j][i] used by j
	std::vector<std::vector<int>> sv((size_t)g_size, std::vector<int>((size_t)g_size, -1));
	for (int i = 0; i < g_size; i++)
		for (int j = i + 1; j < g_size; j++) {
			int s[2];
			if (socketpair(AF_UNIX, SOCK_STREAM, 0, s) != 0) die("socketpair failed");
			int sz = 1 << 20;
			setsockopt(s[0], SOL_SOCKET, SO_SNDBUF, &sz, sizeof sz);
			setsockopt(s[1], SOL_SOCKET, SO_SNDBUF, &sz, sizeof sz);
			sv[(size_t)i][(size_t)j] = s[0];
			sv[(size_t)j][(size_t)i] = s[1];
		}
	fflush(stdout);
	fflush(stderr);
	for (int r = 1; r < g_size; r++) {
		pid_t pid = fork();
		if (pid < 0) die("fork failed");
		if (pid == 0) {
			g_rank = r;
			g_children.clear();
			break;
		}
		g_children.push_back(pid);
	}
	for (int i = 0; i < g_size; i++)
		for (int j = 0; j < g_size; j++) {
			if (i == j) continue;
			int fd = sv[(size_t)i][(size_t)j];
			if (i == g_rank) {
				g_fd[(size_t)j] = fd;
				fcntl(fd, F_SETFL, fcntl(fd, F_GETFL, 0) | O_NONBLOCK);
			} else {
				close(fd);
			}
		}
	signal(SIGPIPE, SIG_IGN);
	return MPI_SUCCESS;
}

int MPI_Initialized(int* flag) {
	*flag = g_init ? 1 : 0;
	return MPI_SUCCESS;
}

int MPI_Finalize(void) {
	if (g_final) return MPI_SUCCESS;
	g_final = true;
	for (FILE* f : g_dump)
		if (f) fclose(f);
	g_dump.clear();
	if (g_size > 1) barrier_impl();
	fflush(stdout);
	fflush(stderr);
	if (g_rank != 0) _exit(0);   // children never return into the caller's atexit chain twice
	for (pid_t c : g_children) {
		int st;
		waitpid(c, &st, 0);
	}
	return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int code) {
	fprintf(stderr, "[mini-mpi rank %d] MPI_Abort(%d)\n", g_rank, code);
	fflush(stderr);
	if (g_sh) g_sh->abort_code.store(code ? code : 1);
	_exit(code ? code : 1);
}

int MPI_Comm_rank(MPI_Comm, int* rank) {
	*rank = g_rank;
	return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm, int* size) {
	*size = g_size;
	return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm) {
	barrier_impl();
	return MPI_SUCCESS;
}
int MPI_Get_processor_name(char* name, int* len) {
	strcpy(name, "localhost");
	*len = 9;
	return MPI_SUCCESS;
}
int MPI_Pcontrol(const int, ...) { return MPI_SUCCESS; }

int MPI_Allgather(const void* sbuf, int scount, MPI_Datatype st, void* rbuf, int, MPI_Datatype, MPI_Comm) {
	size_t n = (size_t)scount * g_types[(size_t)st].extent;
	if (g_size == 1) {
		memcpy(rbuf, sbuf, n);
		return MPI_SUCCESS;
	}
	if (n > SLOT_BYTES) die("collective payload too large");
	memcpy(g_slots + (size_t)g_rank * SLOT_BYTES, sbuf, n);
	barrier_impl();
	for (int r = 0; r < g_size; r++) memcpy((char*)rbuf + (size_t)r * n, g_slots + (size_t)r * SLOT_BYTES, n);
	barrier_impl();
	return MPI_SUCCESS;
}

int MPI_Gather(const void* sbuf, int scount, MPI_Datatype st, void* rbuf, int, MPI_Datatype, int root, MPI_Comm) {
	size_t n = (size_t)scount * g_types[(size_t)st].extent;
	if (g_size == 1) {
		memcpy(rbuf, sbuf, n);
		return MPI_SUCCESS;
	}
	if (n > SLOT_BYTES) die("collective payload too large");
	memcpy(g_slots + (size_t)g_rank * SLOT_BYTES, sbuf, n);
	barrier_impl();
	if (g_rank == root)
		for (int r = 0; r < g_size; r++) memcpy((char*)rbuf + (size_t)r * n, g_slots + (size_t)r * SLOT_BYTES, n);
	barrier_impl();
	return MPI_SUCCESS;
}

int MPI_Allreduce(const void* sbuf, void* rbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm) {
	size_t n = (size_t)count * g_types[(size_t)t].extent;
	if (g_size == 1) {
		memmove(rbuf, sbuf, n);
		return MPI_SUCCESS;
	}
	if (n > SLOT_BYTES) die("collective payload too large");
	memcpy(g_slots + (size_t)g_rank * SLOT_BYTES, sbuf, n);
	barrier_impl();
	// rank-ordered reduction: identical result on every rank
	memcpy(rbuf, g_slots, n);
	for (int r = 1; r < g_size; r++) reduce_any(rbuf, g_slots + (size_t)r * SLOT_BYTES, count, t, op);
	barrier_impl();
	return MPI_SUCCESS;
}

int MPI_Reduce(const void* sbuf, void* rbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm) {
	size_t n = (size_t)count * g_types[(size_t)t].extent;
	if (g_size == 1) {
		memmove(rbuf, sbuf, n);
		return MPI_SUCCESS;
	}
	if (n > SLOT_BYTES) die("collective payload too large");
	memcpy(g_slots + (size_t)g_rank * SLOT_BYTES, sbuf, n);
	barrier_impl();
	if (g_rank == root) {
		memcpy(rbuf, g_slots, n);
		for (int r = 1; r < g_size; r++) reduce_any(rbuf, g_slots + (size_t)r * SLOT_BYTES, count, t, op);
	}
	barrier_impl();
	return MPI_SUCCESS;
}

static int make_req(Kind k, const void* buf, int count, MPI_Datatype t, int peer, int tag, bool persistent) {
	int r = new_req();
	Req& rq = g_reqs[(size_t)r];
	rq.kind = k;
	rq.peer = peer;
	rq.tag = tag;
	rq.buf = (char*)buf;
	rq.nbytes = (size_t)count * g_types[(size_t)t].extent;
	rq.persistent = persistent;
	rq.done = true;
	rq.active = false;
	return r;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm, MPI_Request* req) {
	int r = make_req(SEND, buf, count, t, dest, tag, false);
	if (g_types[(size_t)t].is_fp) dump_payload(dest, tag, buf, g_reqs[(size_t)r].nbytes);
	*req = r;
	post(r);
	return MPI_SUCCESS;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Request* req) {
	int r = make_req(RECV, buf, count, t, src, tag, false);
	*req = r;
	post(r);
	return MPI_SUCCESS;
}
int MPI_Send_init(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm, MPI_Request* req) {
	*req = make_req(SEND, buf, count, t, dest, tag, true);
	return MPI_SUCCESS;
}
int MPI_Recv_init(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Request* req) {
	*req = make_req(RECV, buf, count, t, src, tag, true);
	return MPI_SUCCESS;
}
int MPI_Start(MPI_Request* req) {
	Req& rq = g_reqs[(size_t)*req];
	if (rq.kind == SEND) dump_payload(rq.peer, rq.tag, rq.buf, rq.nbytes);
	post(*req);
	return MPI_SUCCESS;
}
int MPI_Startall(int n, MPI_Request* reqs) {
	// receives first so that incoming data streams straight into user buffers
	for (int i = 0; i < n; i++)
		if (g_reqs[(size_t)reqs[i]].kind == RECV) MPI_Start(&reqs[i]);
	for (int i = 0; i < n; i++)
		if (g_reqs[(size_t)reqs[i]].kind == SEND) MPI_Start(&reqs[i]);
	return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request* req, MPI_Status*) {
	if (*req == MPI_REQUEST_NULL) return MPI_SUCCESS;
	wait_req(*req);
	finish_req(req);
	return MPI_SUCCESS;
}
int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status*) {
	for (int i = 0; i < n; i++) {
		if (reqs[i] == MPI_REQUEST_NULL) continue;
		if (g_reqs[(size_t)reqs[i]].persistent && !g_reqs[(size_t)reqs[i]].active) continue;
		wait_req(reqs[i]);
		finish_req(&reqs[i]);
	}
	return MPI_SUCCESS;
}
int MPI_Testall(int n, MPI_Request* reqs, int* flag, MPI_Status*) {
	progress_all();
	*flag = 1;
	for (int i = 0; i < n; i++)
		if (reqs[i] != MPI_REQUEST_NULL && !g_reqs[(size_t)reqs[i]].done) *flag = 0;
	if (*flag)
		for (int i = 0; i < n; i++) finish_req(&reqs[i]);
	return MPI_SUCCESS;
}

int MPI_Type_create_struct(int n, const int* blocklens, const MPI_Aint* disps, const MPI_Datatype* types, MPI_Datatype* newtype) {
	size_t end = 0, align = 1;
	for (int i = 0; i < n; i++) {
		size_t e = g_types[(size_t)types[i]].extent;
		size_t hi = (size_t)disps[i] + (size_t)blocklens[i] * e;
		if (hi > end) end = hi;
		size_t a = e > 8 ? 8 : e;
		if (a > align) align = a;
	}
	end = (end + align - 1) / align * align;
	g_types.push_back({end, false});
	*newtype = (int)g_types.size() - 1;
	return MPI_SUCCESS;
}
int MPI_Type_create_resized(MPI_Datatype, MPI_Aint, MPI_Aint extent, MPI_Datatype* newtype) {
	g_types.push_back({(size_t)extent, false});
	*newtype = (int)g_types.size() - 1;
	return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }

#define UNSUPPORTED(name) \
	die(name " is not implemented by the mini-MPI shim")

// (blocking) internal send of a reserved-tag message; incoming traffic keeps being drained meanwhile
static void rma_send(int peer, int32_t tag, const char* data, size_t nbytes) {
	int r = new_req();
	Req& rq = g_reqs[(size_t)r];
	rq.kind = SEND;
	rq.peer = peer;
	rq.tag = tag;
	rq.buf = const_cast<char*>(data);
	rq.nbytes = nbytes;
	post(r);
	wait_req(r);
	g_reqs[(size_t)r].in_use = false;
}
static Window& win_of(MPI_Win win) {
	if (win < 0 || (size_t)win >= g_wins.size() || !g_wins[(size_t)win].base) die("invalid window handle");
	return g_wins[(size_t)win];
}

int MPI_Win_create(void* base, MPI_Aint size, int disp_unit, MPI_Info, MPI_Comm, MPI_Win* win) {
	// collective and called in the same order by every rank: the n-th window of every rank gets handle n
	static int n_created = 0;
	const size_t w = (size_t)n_created++;
	if (w >= g_wins.size()) g_wins.resize(w + 1);
	Window& W = g_wins[w];
	W.base = (char*)base;
	W.size = (size_t)size;
	W.disp_unit = disp_unit;
	if (W.snap.empty()) {
		W.snap.resize((size_t)g_size);
		W.done.assign((size_t)g_size, 0);
	}
	*win = (MPI_Win)w;
	barrier_impl();
	return MPI_SUCCESS;
}
int MPI_Win_free(MPI_Win* win) {
	barrier_impl();
	*win = -1;
	return MPI_SUCCESS;
}
int MPI_Win_post(MPI_Group g, int, MPI_Win win) {
	Window& W = win_of(win);
	if (!W.exposed_to.empty()) die("MPI_Win_post on a window whose exposure epoch is still open");
	W.exposed_to = g_groups[(size_t)g];
	for (int r : W.exposed_to) rma_send(r, RMA_SNAP | (int32_t)win, W.base, W.size);
	return MPI_SUCCESS;
}
int MPI_Win_start(MPI_Group g, int, MPI_Win win) {
	Window& W = win_of(win);
	if (!W.accessing.empty()) die("MPI_Win_start on a window whose access epoch is still open");
	W.accessing = g_groups[(size_t)g];
	return MPI_SUCCESS;
}
static const std::vector<char>& rma_copy_of(Window& W, int target) {
	bool member = false;
	for (int r : W.accessing) member |= r == target;
	if (!member) die("MPI_Get outside an access epoch that includes the target");
	while (W.snap[(size_t)target].empty()) {   // the target has not posted yet
		if (!progress_all()) idle_wait();
	}
	return W.snap[(size_t)target].front();
}
int MPI_Get(void* o, int oc, MPI_Datatype ot, int rank, MPI_Aint disp, int tc, MPI_Datatype tt, MPI_Win win) {
	Window& W = win_of(win);
	const size_t nbytes = (size_t)oc * g_types[(size_t)ot].extent;
	if (nbytes != (size_t)tc * g_types[(size_t)tt].extent) die("MPI_Get: origin and target sizes differ");
	const std::vector<char>& copy = rma_copy_of(W, rank);
	const size_t off = (size_t)disp * (size_t)W.disp_unit;   // (the reference creates its windows with disp_unit 1 on every rank)
	if (off + nbytes > copy.size()) die("MPI_Get beyond the end of the target window");
	memcpy(o, copy.data() + off, nbytes);
	return MPI_SUCCESS;
}
int MPI_Rget(void* o, int oc, MPI_Datatype ot, int rank, MPI_Aint disp, int tc, MPI_Datatype tt, MPI_Win win, MPI_Request* req) {
	MPI_Get(o, oc, ot, rank, disp, tc, tt, win);
	*req = MPI_REQUEST_NULL;   // complete at once
	return MPI_SUCCESS;
}
int MPI_Win_complete(MPI_Win win) {
	Window& W = win_of(win);
	for (int r : W.accessing) {
		rma_copy_of(W, r);   // an epoch without a get towards r still consumes r's exposure
		W.snap[(size_t)r].pop_front();
		rma_send(r, RMA_DONE | (int32_t)win, nullptr, 0);
	}
	W.accessing.clear();
	return MPI_SUCCESS;
}
int MPI_Win_wait(MPI_Win win) {
	Window& W = win_of(win);
	for (int r : W.exposed_to) {
		while (W.done[(size_t)r] == 0) {
			if (!progress_all()) idle_wait();
		}
		W.done[(size_t)r]--;
	}
	W.exposed_to.clear();
	return MPI_SUCCESS;
}
int MPI_Win_lock(int, int, int, MPI_Win) { UNSUPPORTED("MPI_Win_lock (passive target synchronisation)"); }
int MPI_Win_lock_all(int, MPI_Win) { UNSUPPORTED("MPI_Win_lock_all (passive target synchronisation)"); }
int MPI_Comm_group(MPI_Comm, MPI_Group* g) {
	g_groups[0].resize((size_t)g_size);
	for (int r = 0; r < g_size; r++) g_groups[0][(size_t)r] = r;
	*g = 0;
	return MPI_SUCCESS;
}
int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* out) {
	std::vector<int> members;
	for (int i = 0; i < n; i++) members.push_back(g_groups[(size_t)g][(size_t)ranks[i]]);
	g_groups.push_back(members);
	*out = (MPI_Group)g_groups.size() - 1;
	return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group*) { return MPI_SUCCESS; }
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm*) { UNSUPPORTED("MPI_Comm_split"); }
int MPI_Dist_graph_create_adjacent(MPI_Comm, int, const int*, const int*, int, const int*, const int*, MPI_Info, int, MPI_Comm*) {
	UNSUPPORTED("MPI_Dist_graph_create_adjacent");
}
int MPI_Dist_graph_neighbors(MPI_Comm, int, int*, int*, int, int*, int*) { UNSUPPORTED("MPI_Dist_graph_neighbors"); }

}  // extern "C"
