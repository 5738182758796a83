"""ctypes binding of ``libdrgnn.so`` (C-ABI declared in ``include/drgnn.h``).

There is no CPU fallback: if the library is missing or does not export a symbol the
import of anything that computes raises.  Structures mirror the header field by field.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libdrgnn.so')

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)
VP = C.c_void_p

ST_EDGE_OUTSIDE_GRAPH = 1
ST_CLUSTER_RANGE = 2
ST_CLUSTER1_LENGTH = 4
ST_CLUSTER_ORDER = 8
ST_NEGATIVE_ID = 16
ST_FUSED_BOUNDS = 64
STATUS_TEXT = {
    ST_EDGE_OUTSIDE_GRAPH: 'an edge endpoint lies outside its graph',
    ST_CLUSTER_RANGE: 'cluster-id range of one graph exceeds 32768',
    ST_CLUSTER1_LENGTH: 'len(cluster1) of a graph differs from its number of level-0 clusters',
    ST_CLUSTER_ORDER: 'global cluster ids do not increase with the graph id',
    ST_NEGATIVE_ID: 'negative cluster id',
    ST_FUSED_BOUNDS: 'a graph exceeds the per-graph bounds given to the fused kernels',
    128: 'the grid barrier of the fused gradient reduction timed out (the step kernel was not co-resident)',
}


class StructureIO(C.Structure):
    _fields_ = [
        ('B', C.c_int32), ('N', C.c_int32), ('E', C.c_int32), ('L1', C.c_int32), ('ne', C.c_int32),
        ('max_n', C.c_int32), ('max_e', C.c_int32), ('clusters_are_local', C.c_int32),
        ('idx32', C.c_int32), ('edge16', C.c_int32),
        ('node_ptr', VP), ('edge_ptr', VP), ('c1_ptr', VP), ('edge_index', VP), ('edge_attr', VP),
        ('cluster0', VP), ('cluster1', VP),
        ('rowptr0', VP), ('col0', VP), ('eid0', VP), ('cscptr0', VP), ('cscrow0', VP), ('csceid0', VP),
        ('w0csr', VP), ('w0csc', VP),
        ('cl0', VP), ('cl0_i64', VP), ('cmptr0', VP), ('cmem0', VP), ('kptr0', VP), ('batch1', VP),
        ('batch1_i64', VP),
        ('rowptr1', VP), ('col1', VP), ('edge_index1', VP), ('edge_attr1', VP), ('cscptr1', VP),
        ('cscrow1', VP), ('csceid1', VP), ('w1csc', VP),
        ('cl1', VP), ('cmptr1', VP), ('cmem1', VP), ('kptr1', VP), ('batch2', VP), ('batch2_i64', VP),
        ('counts', VP), ('status', VP),
        ('gstat', VP), ('scratch_n', VP), ('scratch_e', VP), ('scratch_f', VP),
        ('blob', VP), ('wblob', VP),
        ('x', VP), ('zin1', VP), ('F', C.c_int32), ('ld_zin1', C.c_int32), ('zin_kind', C.c_int32),
        ('launch_flags', C.c_int32), ('max_k', C.c_int32), ('max_q', C.c_int32),
    ]


class AggregateArgs(C.Structure):
    _fields_ = [
        ('src', VP), ('ld_src', C.c_int32),
        ('out', VP), ('ld_out', C.c_int32),
        ('rowptr', VP), ('col', VP),
        ('ew', VP), ('sscale', VP),
        ('self_src', VP), ('ld_self', C.c_int32),
        ('self_out', VP), ('ld_self_out', C.c_int32),
        ('selfc_in', VP), ('selfc_out', VP),
        ('post_out', VP),
        ('bias', VP),
        ('n_rows', C.c_int32), ('n_rows_dev', VP),
        ('C', C.c_int32), ('post_mode', C.c_int32), ('self_mode', C.c_int32), ('relu', C.c_int32),
    ]


class LinearArgs(C.Structure):
    _fields_ = [
        ('X', VP), ('ldx', C.c_int32),
        ('W', VP), ('bias', VP),
        ('Y', VP), ('ldy', C.c_int32),
        ('out_mask', VP), ('ld_mask', C.c_int32), ('mask_scale', C.c_float),
        ('rows', C.c_int32), ('rows_dev', VP),
        ('Fin', C.c_int32), ('Fout', C.c_int32), ('groups', C.c_int32),
        ('w_layout', C.c_int32), ('relu', C.c_int32), ('math', C.c_int32),
    ]


class LinearWgradArgs(C.Structure):
    _fields_ = [
        ('X', VP), ('ldx', C.c_int32),
        ('G', VP), ('ldg', C.c_int32),
        ('dW', VP), ('dbias', VP),
        ('rows', C.c_int32), ('rows_dev', VP),
        ('Fin', C.c_int32), ('Fout', C.c_int32), ('groups', C.c_int32),
        ('w_layout', C.c_int32), ('accumulate', C.c_int32),
        ('work', VP), ('work_floats', C.c_int64),
    ]


class GinetFusedArgs(C.Structure):
    _fields_ = [
        ('B', C.c_int32), ('F', C.c_int32), ('h1', C.c_int32), ('h2', C.c_int32), ('nb', C.c_int32),
        ('max_n', C.c_int32), ('max_k', C.c_int32), ('max_q', C.c_int32),
        ('node_ptr', VP),
        ('rowptr0', VP), ('col0', VP),
        ('rowptr1', VP), ('col1', VP),
        ('cscptr1', VP), ('cscrow1', VP),
        ('cmptr0', VP), ('cmem0', VP), ('cl0', VP), ('kptr0', VP),
        ('cmptr1', VP), ('cmem1', VP), ('cl1', VP), ('kptr1', VP),
        ('status', VP),
        ('W1', VP), ('W2', VP),
        ('x', VP),
        ('Zin1', VP), ('Z1', VP), ('arg0', VP),
        ('Zin2', VP), ('Z2', VP), ('arg1', VP),
        ('R', VP),
        ('dR', VP), ('partial', VP), ('dW1', VP), ('dW2', VP),
    ]


class GinetStepArgs(C.Structure):
    _fields_ = [
        ('g', GinetFusedArgs),
        ('fc1_w', VP), ('fc1_b', VP), ('fc2_w', VP), ('fc2_b', VP),
        ('Hd', C.c_int32), ('out', C.c_int32),
        ('keep', VP), ('keep_scale', C.c_float),
        ('y', VP), ('y_class', VP), ('class_w', VP),
        ('task', C.c_int32), ('inv_norm', C.c_float),
        ('pred', VP), ('loss', VP),
        ('partial', VP), ('partial_ld', C.c_int64),
        ('grads', VP), ('n_params', C.c_int32),
        ('off_w1', C.c_int32), ('off_w2', C.c_int32), ('off_fc1w', C.c_int32), ('off_fc1b', C.c_int32),
        ('off_fc2w', C.c_int32), ('off_fc2b', C.c_int32),
        ('forward_only', C.c_int32),
        ('head_off', C.c_int32),
        ('drop_p', C.c_float), ('seed', C.c_uint32),
        ('fuse_adam', C.c_int32), ('lr', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
        ('adam_p', VP), ('adam_m', VP), ('adam_v', VP), ('step_dev', VP),
        ('skip_reduce', C.c_int32), ('flags', C.c_int32), ('max_e', C.c_int32), ('variant', C.c_int32),
        ('blob', VP), ('edge_ptr', VP), ('comm', VP), ('gdesc', VP), ('zin1', VP),
    ]


class NetStepArgs(C.Structure):
    _fields_ = [
        ('kind', C.c_int32), ('B', C.c_int32), ('F', C.c_int32), ('h1', C.c_int32), ('h2', C.c_int32), ('Hd', C.c_int32),
        ('out', C.c_int32),
        ('max_n', C.c_int32), ('max_e', C.c_int32), ('max_k', C.c_int32), ('max_q', C.c_int32),
        ('tiles', C.c_int32),
        ('x', VP),
        ('blob', VP), ('wblob', VP), ('gdesc', VP),
        ('node_ptr', VP), ('edge_ptr', VP),
        ('params', VP),
        ('off_w1', C.c_int32), ('off_b1', C.c_int32), ('off_w2', C.c_int32), ('off_b2', C.c_int32),
        ('off_fc1w', C.c_int32), ('off_fc1b', C.c_int32), ('off_fc2w', C.c_int32), ('off_fc2b', C.c_int32),
        ('keep', VP), ('keep_scale', C.c_float), ('drop_p', C.c_float), ('seed', C.c_uint32),
        ('y', VP), ('y_class', VP), ('class_w', VP),
        ('task', C.c_int32), ('inv_norm', C.c_float), ('forward_only', C.c_int32), ('skip_reduce', C.c_int32),
        ('pred', VP), ('loss', VP), ('R', VP),
        ('partial', VP), ('partial_ld', C.c_int64),
        ('grads', VP), ('n_params', C.c_int32),
        ('fuse_adam', C.c_int32), ('lr', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
        ('flags', C.c_int32),
        ('adam_p', VP), ('adam_m', VP), ('adam_v', VP), ('step_dev', VP),
        ('status', VP),
        ('comm', VP),
        ('kptr0', VP), ('kptr1', VP),
        ('Zin1', VP), ('Z1', VP), ('arg0', VP), ('Zin2', VP), ('Z2', VP), ('arg1', VP),
        ('layers3', C.c_int32), ('off_w3', C.c_int32), ('off_b3', C.c_int32), ('reserved3', C.c_int32),
        ('zin1', VP),
    ]


class FeedStep(C.Structure):
    _fields_ = [
        ('h_src', VP), ('d_dst', VP), ('nbytes', C.c_int64),
        ('prep_graph', VP), ('step_graph', VP),
        ('d_out', VP), ('h_out', VP), ('out_bytes', C.c_int64),
        ('slot', C.c_int32), ('reserved', C.c_int32),
    ]


MAX_PEERS = 8
IPC_HANDLE_BYTES = 64
NCCL_ID_BYTES = 128


class PeerComm(C.Structure):
    _fields_ = [
        ('world', C.c_int32), ('rank', C.c_int32),
        ('xbuf', VP * MAX_PEERS), ('xflag', VP * MAX_PEERS),
        ('ctr', VP), ('stride', C.c_int64), ('max_blocks', C.c_int32), ('reserved', C.c_int32),
        ('timeout_ns', C.c_uint64),
        ('xll', VP * MAX_PEERS),
    ]


class PeerAdamArgs(C.Structure):
    _fields_ = [
        ('partial', VP), ('B', C.c_int32), ('reserved0', C.c_int32), ('partial_ld', C.c_int64),
        ('grads', VP), ('n_params', C.c_int32), ('n_sum', C.c_int32),
        ('apply_adam', C.c_int32), ('lr', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
        ('reserved1', C.c_int32),
        ('adam_p', VP), ('adam_m', VP), ('adam_v', VP), ('step_dev', VP),
    ]


class HeadArgs(C.Structure):
    _fields_ = [
        ('R', VP), ('ldr', C.c_int32),
        ('W1', VP), ('b1', VP),
        ('W2', VP), ('b2', VP),
        ('keep', VP), ('keep_scale', C.c_float),
        ('y', VP), ('y_class', VP), ('class_w', VP),
        ('B', C.c_int32), ('C', C.c_int32), ('Hd', C.c_int32), ('out', C.c_int32),
        ('task', C.c_int32), ('inv_norm', C.c_float),
        ('pred', VP), ('loss', VP), ('H', VP),
        ('dW1', VP), ('db1', VP), ('dW2', VP), ('db2', VP),
        ('dR', VP), ('lddr', C.c_int32),
    ]


_i32, _i64, _f32 = C.c_int32, C.c_int64, C.c_float
_SIGNATURES = {
    'drgnn_last_error': (C.c_char_p, []),
    'drgnn_version': (C.c_int, []),
    'drgnn_device_sms': (C.c_int, []),
    'drgnn_device_smem_optin': (C.c_int, []),
    'drgnn_structure_smem_bytes': (_i64, [_i32, _i32, _i32]),
    'drgnn_structure_build': (C.c_int, [C.POINTER(StructureIO), VP]),
    'drgnn_cluster_offset': (C.c_int, [VP, VP, _i32, VP, VP]),
    'drgnn_ptr_from_sorted_ids': (C.c_int, [VP, _i32, _i32, VP, VP, VP]),
    'drgnn_aggregate': (C.c_int, [C.POINTER(AggregateArgs), VP]),
    'drgnn_aggregate_tiled': (C.c_int, [C.POINTER(AggregateArgs), VP, VP, _i32, _i32, _i32, VP]),
    'drgnn_linear': (C.c_int, [C.POINTER(LinearArgs), VP]),
    'drgnn_linear_tcgen05_supported': (C.c_int, [C.POINTER(LinearArgs)]),
    'drgnn_linear_tcgen05': (C.c_int, [C.POINTER(LinearArgs), VP]),
    'drgnn_debug_tc5_cycles': (C.c_int, [C.POINTER(C.c_uint64)]),
    'drgnn_linear_wgrad_work_floats': (_i64, [_i32, _i32, _i32, _i32]),
    'drgnn_linear_wgrad': (C.c_int, [C.POINTER(LinearWgradArgs), VP]),
    'drgnn_maxpool_fwd': (C.c_int, [VP, _i32, VP, VP, _i32, VP, _i32, VP, _i32, VP, VP]),
    'drgnn_maxpool_bwd': (C.c_int, [VP, _i32, VP, VP, VP, _i32, _i32, VP, _i32, VP, _i32, VP]),
    'drgnn_segment_mean_fwd': (C.c_int, [VP, _i32, VP, _i32, _i32, VP, _i32, VP]),
    'drgnn_segment_mean_bwd': (C.c_int, [VP, _i32, VP, _i32, _i32, VP, _i32, VP]),
    'drgnn_mse_loss': (C.c_int, [VP, VP, _i32, _f32, _i32, VP, VP, VP]),
    'drgnn_ce_loss': (C.c_int, [VP, _i32, VP, VP, _i32, _i32, _f32, VP, VP, VP]),
    'drgnn_adam_flat': (C.c_int, [VP, VP, VP, VP, VP, _i64, _f32, _f32, _f32, _f32, _f32, VP]),
    'drgnn_ginet_fused_smem_bytes': (_i64, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    'drgnn_ginet_fused_fwd': (C.c_int, [C.POINTER(GinetFusedArgs), VP]),
    'drgnn_ginet_fused_bwd': (C.c_int, [C.POINTER(GinetFusedArgs), VP]),
    'drgnn_ginet_step_smem_bytes': (_i64, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    'drgnn_ginet_step': (C.c_int, [C.POINTER(GinetStepArgs), VP]),
    'drgnn_ginet_step2_smem_bytes': (_i64, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    'drgnn_ginet_step_last_variant': (C.c_int, []),
    'drgnn_ginet_step_last_launches': (C.c_int, []),
    'drgnn_ginet_step2_max_clusters': (C.c_int, [_i64]),
    'drgnn_debug_phase_cycles': (C.c_int, [C.POINTER(C.c_uint64)]),
    'drgnn_debug_blob_cycles': (C.c_int, [C.POINTER(C.c_uint64)]),
    'drgnn_structure_blob_smem_bytes': (_i64, [_i32, _i32]),
    'drgnn_structure_blob_smem_bytes_ex': (_i64, [_i32, _i32, _i32, _i32, _i32, _i32]),
    'drgnn_structure_blob': (C.c_int, [C.POINTER(StructureIO), VP]),
    'drgnn_net_step_smem_bytes': (_i64, [_i32] * 11),
    'drgnn_net_step_pick_tiles': (C.c_int, [_i32] * 10),
    'drgnn_net_step_max_clusters': (C.c_int, [_i32, _i32, _i64]),
    'drgnn_net_step_smem_bytes_l': (_i64, [_i32] * 12),
    'drgnn_net_step_pick_tiles_l': (C.c_int, [_i32] * 11),
    'drgnn_net_step': (C.c_int, [C.POINTER(NetStepArgs), VP]),
    'drgnn_sgat_step': (C.c_int, [C.POINTER(NetStepArgs), VP]),
    'drgnn_fout_step': (C.c_int, [C.POINTER(NetStepArgs), VP]),
    'drgnn_net_step_last_launches': (C.c_int, []),
    'drgnn_net_step_last_tiles': (C.c_int, []),
    'drgnn_debug_phase3_cycles': (C.c_int, [C.POINTER(C.c_uint64)]),
    'drgnn_debug_cta_times': (C.c_int, [C.POINTER(C.c_uint64), C.c_int32]),
    'drgnn_mcl_work_doubles': (_i64, [_i32, _i32]),
    'drgnn_mcl_cluster': (C.c_int, [VP, VP, VP, _i32, _i64, _i32, _i32, VP, VP, VP, VP, VP]),
    'drgnn_feed_run': (C.c_int, [C.POINTER(FeedStep), _i32, _i32, VP, VP, VP, VP, VP, VP, _i64, _i32]),
    'drgnn_debug_structure_cycles': (C.c_int, [C.POINTER(C.c_uint64)]),
    'drgnn_head_smem_bytes': (_i64, [_i32, _i32, _i32]),
    'drgnn_head': (C.c_int, [C.POINTER(HeadArgs), VP]),
    'drgnn_relu_mask': (C.c_int, [VP, _i32, VP, _i32, _i32, VP, _i32, VP, _i32, VP]),
    'drgnn_fill_f32': (C.c_int, [VP, _f32, _i64, VP]),
    'drgnn_fill_i32': (C.c_int, [VP, _i32, _i64, VP]),
    'drgnn_comm_alloc': (C.c_int, [_i64, C.POINTER(VP), C.c_char_p]),
    'drgnn_comm_open': (C.c_int, [C.c_char_p, C.POINTER(VP)]),
    'drgnn_comm_close': (C.c_int, [VP]),
    'drgnn_comm_free': (C.c_int, [VP]),
    'drgnn_comm_status': (C.c_int, [VP, C.POINTER(C.c_uint32)]),
    'drgnn_peer_reduce_adam': (C.c_int, [C.POINTER(PeerComm), C.POINTER(PeerAdamArgs), VP]),
    'drgnn_nccl_available': (C.c_int, []),
    'drgnn_nccl_unique_id': (C.c_int, [VP]),
    'drgnn_nccl_init': (C.c_int, [C.POINTER(VP), _i32, _i32, VP]),
    'drgnn_nccl_allreduce': (C.c_int, [VP, VP, _i64, VP]),
    'drgnn_nccl_destroy': (C.c_int, [VP]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))

_lib = None
launch_count = 0           # C-ABI compute calls made by this process
kernel_count = 0           # CUDA kernels those calls launched (bench: gpu_launches)
KERNELS_PER_CALL = {'drgnn_structure_build': 2, 'drgnn_cluster_offset': 2, 'drgnn_ginet_fused_bwd': 2, 'drgnn_ginet_step': 2,
                    'drgnn_adam_flat': 2}   # lower bounds for the others


class DrgnnError(RuntimeError):
    pass


def load():
    """Load libdrgnn.so (once).  Raises if it has not been built - there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DrgnnError('libdrgnn.so is missing (%s): build it with `python -m deeprank_gnn_b200.build` '
                         'or __graft_entry__.build(); this package has no CPU / eager fallback' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().drgnn_last_error()
        raise DrgnnError('%s failed (code %d): %s' % (what, rc, msg.decode() if msg else '?'))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DrgnnError('the hot path runs on CUDA only (got a %s tensor); there is no CPU fallback' % t.device)


def call(name, *args):
    """Invoke one C-ABI entry point, raising DrgnnError on a non-zero status."""
    global launch_count, kernel_count
    rc = getattr(load(), name)(*args)
    launch_count += 1
    kernel_count += KERNELS_PER_CALL.get(name, 1)
    check(rc, name)
