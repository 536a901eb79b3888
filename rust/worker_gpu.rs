//! GPU worker: `src/worker.rs` of MengLing-L/distributed_plonk with every hot-path body replaced by ONE call into
//! libdplonk.so (include/dplonk.h).  The Cap'n Proto surface (hello_world.capnp), the dispatcher, config/network.json
//! and the two listeners are the reference's; what changes is where the arithmetic runs:
//!
//!   reference body (worker.rs)                          here
//!   init          126-157  store bases, 6 domains        dp_init (+ one-time peer-arena handshake, see below)
//!   var_msm       159-185  VariableBaseMSM               dp_msm
//!   fft_init      187-233  FftTask{rows, cols}           dp_fft_init
//!   fft1          235-278  fft1_helper on one row        dp_fft1
//!   fft2_prepare  280-345  W TCP connects + fftExchange  dp_fft2_prepare: the row kernel stores into the peers' GPUs
//!   fft_exchange  412-438  scatter into cols             (not sent any more; carries the arena handles at start-up)
//!   fft2          347-381  fft2_helper per column        dp_fft2
//!   round1        383-408  ifft, blind, commit           dp_round1 (blinders still drawn from the worker's ThreadRng)
//!
//! NOT COMPILED in this repository (no Rust toolchain in the build image); the same call sequences are exercised
//! through the C ABI by distributed_plonk_b200/worker.py and tests/.  Build: see rust/README.md and rust/build.rs.
//!
//! Peer transport.  Workers of one multi-GPU box exchange the row-phase output through CUDA-IPC arenas
//! (dp_peer_arena_create / dp_peer_attach): the 64-byte handles travel once, at the first `init`, over the existing
//! PlonkPeer.fftExchange RPC with the reserved task id u64::MAX (`from` = sender, `v` = [handle]) - the schema is
//! untouched.  After that `fft2Prepare` needs no peer RPC at all: it returns when this worker's stores into the
//! owners' receive matrices are complete, and the dispatcher's join over all fft2Prepare replies
//! (dispatcher2.rs:767-772) is the barrier before `fft2`.  The arena holds ARENA_SLOTS receive matrices: exchange
//! number k lands in slot k mod ARENA_SLOTS on every worker, which is consistent because the dispatcher's single
//! thread writes each task's fft2Prepare to all connections at once (every worker sees the same order).  The
//! dispatcher of the reference keeps at most 26 transforms in flight (join_all over 25 + 1, dispatcher2.rs:382-414);
//! 32 slots of 2^25 * 32 / W bytes are 16 / 8 / 4 GiB per worker at W = 2 / 4 / 8.
#![feature(int_roundings)]

use ark_bls12_381::Fr;
use ark_std::UniformRand;
use capnp::{capability::Promise, message::ReaderOptions};
use capnp_rpc::{rpc_twoparty_capnp, twoparty, RpcSystem};
use futures::AsyncReadExt;
use hello_world::{
    config::NetworkConfig,
    dplonk_sys::*,
    hello_world_capnp::{plonk_peer, plonk_slave},
    utils::serialize,
};
use rand::{rngs::ThreadRng, thread_rng};
use std::{
    collections::HashMap,
    fs::File,
    ptr,
    sync::Arc,
};

const FR_BYTES: usize = 32; // size_of::<Fr>()            (utils.rs:27-43 raw structs)
const G1_AFFINE_BYTES: usize = 104; // size_of::<G1Affine>()
const G1_PROJECTIVE_BYTES: usize = 144; // size_of::<G1Projective>()
const HANDSHAKE_ID: u64 = u64::MAX;
const ARENA_SLOTS: usize = 32;

struct State {
    rng: ThreadRng,
    me: usize,
    network: NetworkConfig,
    ctx: *mut dp_ctx,
    /// (columns of this worker, r) per open task: the shape of the fft2 reply
    dims: HashMap<u64, (usize, usize)>,
    r: [usize; 2], // r of the gate / quotient domain (worker.rs:144-154)
    c: [usize; 2],
    /// fused exchange: handles that arrived before this worker created its own arena
    early_handles: Vec<(u64, Vec<u8>)>,
    arena_ready: bool,
}

#[derive(Clone)]
struct PlonkImpl {
    state: Arc<State>,
}

impl PlonkImpl {
    #[allow(clippy::mut_from_ref)]
    fn st(&self) -> &mut State {
        // single-threaded LocalSet, as in the reference (worker.rs:441,453), which does the same cast
        unsafe { &mut *(Arc::as_ptr(&self.state) as *mut State) }
    }
}

fn concat(list: capnp::data_list::Reader) -> Vec<u8> {
    let mut out = vec![];
    list.iter().for_each(|chunk| out.extend_from_slice(chunk.unwrap()));
    out
}

/// send this worker's arena handle to every peer over PlonkPeer.fftExchange(id = HANDSHAKE_ID)
async fn announce_handle(network: &NetworkConfig, me: usize, handle: Vec<u8>) {
    for (q, peer) in network.peers.iter().enumerate() {
        if q == me {
            continue;
        }
        let stream = tokio::net::TcpStream::connect(peer).await.unwrap();
        stream.set_nodelay(true).unwrap();
        let (reader, writer) = tokio_util::compat::TokioAsyncReadCompatExt::compat(stream).split();
        let mut rpc_system = RpcSystem::new(
            Box::new(twoparty::VatNetwork::new(reader, writer, rpc_twoparty_capnp::Side::Client, ReaderOptions::new())),
            None,
        );
        let connection = rpc_system.bootstrap::<plonk_peer::Client>(rpc_twoparty_capnp::Side::Server);
        tokio::task::spawn_local(rpc_system);
        let mut request = connection.fft_exchange_request();
        let mut r = request.get();
        r.set_id(HANDSHAKE_ID);
        r.set_from(me as u64);
        r.init_v(1).set(0, &handle);
        request.send().promise.await.unwrap();
    }
}

impl plonk_slave::Server for PlonkImpl {
    fn init(&mut self, params: plonk_slave::InitParams, _: plonk_slave::InitResults) -> Promise<(), capnp::Error> {
        let p = params.get().unwrap();
        let (domain_size, quot_domain_size) = (p.get_domain_size(), p.get_quot_domain_size());
        let bases = concat(p.get_bases().unwrap()); // chunks are cut at 2^28 B irrespective of struct boundaries
        let st = self.st();
        let rc = unsafe { dp_init(st.ctx, bases.as_ptr(), bases.len() / G1_AFFINE_BYTES, domain_size, quot_domain_size) };
        if let Err(e) = check(st.ctx, rc) {
            return Promise::err(e);
        }
        for (k, size) in [domain_size, quot_domain_size].iter().enumerate() {
            let log = (*size as usize).next_power_of_two().trailing_zeros();
            st.r[k] = 1 << (log >> 1);
            st.c[k] = (1usize << log) / st.r[k];
        }
        let n_workers = st.network.peers.len();
        if n_workers == 1 || st.arena_ready {
            return Promise::ok(());
        }
        // one-time: receive arena of ARENA_SLOTS matrices of r*c/W Fr each (the larger domain), handle to every peer
        let slot = (st.r[1] * st.c[1]).max(st.r[0] * st.c[0]) / n_workers * FR_BYTES;
        let mut handle = vec![0u8; DP_IPC_HANDLE_BYTES];
        let rc = unsafe { dp_peer_arena_create(st.ctx, (ARENA_SLOTS * slot) as u64, handle.as_mut_ptr()) };
        if let Err(e) = check(st.ctx, rc) {
            return Promise::err(e);
        }
        st.arena_ready = true;
        for (from, h) in st.early_handles.drain(..) {
            let rc = unsafe { dp_peer_attach(st.ctx, from, h.as_ptr()) };
            if let Err(e) = check(st.ctx, rc) {
                return Promise::err(e);
            }
        }
        let (network, me) = (st.network.clone(), st.me);
        Promise::from_future(async move {
            announce_handle(&network, me, handle).await;
            Ok(())
        })
    }

    fn var_msm(&mut self, params: plonk_slave::VarMsmParams, mut results: plonk_slave::VarMsmResults) -> Promise<(), capnp::Error> {
        let p = params.get().unwrap();
        let workload = p.get_workload().unwrap();
        let scalars = concat(p.get_scalars().unwrap()); // BigInteger256, 32 B each
        let st = self.st();
        let mut out = [0u8; G1_PROJECTIVE_BYTES];
        let rc = unsafe {
            dp_msm(st.ctx, workload.get_start(), workload.get_end(), scalars.as_ptr(), scalars.len() / FR_BYTES, out.as_mut_ptr())
        };
        if let Err(e) = check(st.ctx, rc) {
            return Promise::err(e);
        }
        results.get().set_result(&out); // the normalised GroupProjective (Z = 1) or ark's identity (0, 1, 0)
        Promise::ok(())
    }

    fn fft_init(&mut self, params: plonk_slave::FftInitParams, _: plonk_slave::FftInitResults) -> Promise<(), capnp::Error> {
        let p = params.get().unwrap();
        let wl = p
            .get_workloads()
            .unwrap()
            .into_iter()
            .map(|w| dp_fft_workload {
                row_start: w.get_row_start(),
                row_end: w.get_row_end(),
                col_start: w.get_col_start(),
                col_end: w.get_col_end(),
            })
            .collect::<Vec<_>>();
        let st = self.st();
        let k = p.get_is_quot() as usize;
        let mine = wl[st.me];
        st.dims.insert(p.get_id(), ((mine.col_end - mine.col_start) as usize, st.r[k]));
        let rc = unsafe {
            dp_fft_init(st.ctx, p.get_id(), wl.as_ptr(), wl.len(), p.get_is_quot() as i32, p.get_is_inv() as i32, p.get_is_coset() as i32)
        };
        match check(st.ctx, rc) {
            Ok(()) => Promise::ok(()),
            Err(e) => Promise::err(e),
        }
    }

    fn fft1(&mut self, params: plonk_slave::Fft1Params, _: plonk_slave::Fft1Results) -> Promise<(), capnp::Error> {
        let p = params.get().unwrap();
        let v = concat(p.get_v().unwrap());
        let st = self.st();
        // asynchronous copy-in; the row transform itself runs in fft2_prepare, batched over all rows
        let rc = unsafe { dp_fft1(st.ctx, p.get_id(), p.get_i(), v.as_ptr(), v.len() / FR_BYTES) };
        match check(st.ctx, rc) {
            Ok(()) => Promise::ok(()),
            Err(e) => Promise::err(e),
        }
    }

    fn fft2_prepare(&mut self, params: plonk_slave::Fft2PrepareParams, _: plonk_slave::Fft2PrepareResults) -> Promise<(), capnp::Error> {
        let id = params.get().unwrap().get_id();
        let st = self.st();
        // one worker: the whole transform is queued; several: the row kernels store into the peers' arenas and the call
        // returns when they are done (DP_E_STATE if all ARENA_SLOTS slots are waiting for their fft2)
        let rc = unsafe { dp_fft2_prepare(st.ctx, id) };
        match check(st.ctx, rc) {
            Ok(()) => Promise::ok(()),
            Err(e) => Promise::err(e),
        }
    }

    fn fft2(&mut self, params: plonk_slave::Fft2Params, mut results: plonk_slave::Fft2Results) -> Promise<(), capnp::Error> {
        let id = params.get().unwrap().get_id();
        let st = self.st();
        let (n_cols, r) = match st.dims.remove(&id) {
            Some(d) => d,
            None => return Promise::err(capnp::Error::failed(format!("fft2: unknown task {}", id))),
        };
        let mut buf = vec![0u8; n_cols * r * FR_BYTES];
        let rc = unsafe { dp_fft2(st.ctx, id, buf.as_mut_ptr(), buf.len()) }; // column kernels + copy-out; frees the slot
        if let Err(e) = check(st.ctx, rc) {
            return Promise::err(e);
        }
        let mut builder = results.get().init_v(n_cols as u32);
        for (k, col) in buf.chunks(r * FR_BYTES).enumerate() {
            builder.set(k as u32, col); // one Data per column, as worker.rs:365,375
        }
        Promise::ok(())
    }

    fn round1(&mut self, params: plonk_slave::Round1Params, mut results: plonk_slave::Round1Results) -> Promise<(), capnp::Error> {
        let evals = concat(params.get().unwrap().get_w().unwrap());
        let st = self.st();
        // DensePolynomial::rand(1, rng) of worker.rs:400 draws two coefficients: the blinders stay the worker's secret
        let blind = [Fr::rand(&mut st.rng), Fr::rand(&mut st.rng)];
        let mut out = [0u8; G1_PROJECTIVE_BYTES];
        let rc = unsafe { dp_round1(st.ctx, evals.as_ptr(), evals.len() / FR_BYTES, serialize(&blind).as_ptr(), out.as_mut_ptr()) };
        if let Err(e) = check(st.ctx, rc) {
            return Promise::err(e);
        }
        results.get().set_c(&out);
        Promise::ok(())
    }
}

impl plonk_peer::Server for PlonkImpl {
    fn fft_exchange(&mut self, params: plonk_peer::FftExchangeParams, _: plonk_peer::FftExchangeResults) -> Promise<(), capnp::Error> {
        let p = params.get().unwrap();
        if p.get_id() != HANDSHAKE_ID {
            return Promise::err(capnp::Error::failed(
                "fftExchange: GPU workers exchange through peer memory; a CPU reference worker cannot be mixed in".to_string(),
            ));
        }
        let handle = concat(p.get_v().unwrap());
        let st = self.st();
        if !st.arena_ready {
            st.early_handles.push((p.get_from(), handle));
            return Promise::ok(());
        }
        let rc = unsafe { dp_peer_attach(st.ctx, p.get_from(), handle.as_ptr()) };
        match check(st.ctx, rc) {
            Ok(()) => Promise::ok(()),
            Err(e) => Promise::err(e),
        }
    }
}

async fn serve<C, F>(addr: std::net::SocketAddr, make: F)
where
    C: capnp::capability::FromClientHook + Clone + 'static,
    F: Fn() -> C,
{
    let listener = tokio::net::TcpListener::bind(addr).await.unwrap();
    loop {
        let (stream, _) = listener.accept().await.unwrap();
        stream.set_nodelay(true).unwrap();
        let (reader, writer) = tokio_util::compat::TokioAsyncReadCompatExt::compat(stream).split();
        let network = twoparty::VatNetwork::new(
            reader,
            writer,
            rpc_twoparty_capnp::Side::Server,
            ReaderOptions { traversal_limit_in_words: Some(usize::MAX), nesting_limit: 64 },
        );
        let client: C = make();
        tokio::task::spawn_local(RpcSystem::new(Box::new(network), Some(client.as_client_hook().add_ref().into())));
    }
}

#[tokio::main(flavor = "current_thread")]
pub async fn main() -> Result<(), Box<dyn std::error::Error>> {
    let args = std::env::args().collect::<Vec<_>>();
    if args.len() != 2 {
        println!("usage: {} <me>", args[0]);
        return Ok(());
    }
    let me: usize = args[1].parse().unwrap();
    let network: NetworkConfig = serde_json::from_reader(File::open("config/network.json")?)?;
    // one GPU per worker of the box: worker i on device i unless DPLONK_DEVICE says otherwise
    let device = std::env::var("DPLONK_DEVICE").ok().and_then(|d| d.parse().ok()).unwrap_or(me as i32);
    let mut ctx: *mut dp_ctx = ptr::null_mut();
    let rc = unsafe { dp_create(device, me as u64, network.slaves.len() as u64, &mut ctx) };
    check(ctx, rc).map_err(|e| Box::new(e) as Box<dyn std::error::Error>)?;

    let state = Arc::new(State {
        rng: thread_rng(),
        me,
        network,
        ctx,
        dims: HashMap::new(),
        r: [1, 1],
        c: [1, 1],
        early_handles: vec![],
        arena_ready: false,
    });
    let local = tokio::task::LocalSet::new();
    let s = state.clone();
    local.spawn_local(async move {
        let imp = PlonkImpl { state: s.clone() };
        serve(s.network.slaves[me], move || capnp_rpc::new_client::<plonk_slave::Client, _>(imp.clone())).await;
    });
    let s = state.clone();
    local.spawn_local(async move {
        let imp = PlonkImpl { state: s.clone() };
        serve(s.network.peers[me], move || capnp_rpc::new_client::<plonk_peer::Client, _>(imp.clone())).await;
    });
    local.await;
    unsafe { dp_destroy(ctx) };
    Ok(())
}
