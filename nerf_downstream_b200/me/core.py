"""ME-compatible tensors and coordinate manager.

Mirrors the part of MinkowskiEngine 0.5.x's Python surface the reference uses:
`CoordinateMapKey`, `CoordinateManager`, `KernelGenerator`, `SparseTensor`, `TensorField`
(co3d_3d/src/models/mink/modules/sparse_conv.py:7-12, base_model.py:10-13, res16unet.py:392-435,
layernorm.py:16-30).  All device work goes through `nerf_downstream_b200.ops` (C-ABI kernels).
"""
from __future__ import annotations

import warnings
from enum import Enum
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from .. import lib as L
from .. import ops


# ---------------------------------------------------------------------------
# enums (appendix A.1)
# ---------------------------------------------------------------------------
class SparseTensorQuantizationMode(Enum):
    RANDOM_SUBSAMPLE = 0
    UNWEIGHTED_AVERAGE = 1
    UNWEIGHTED_SUM = 2
    NO_QUANTIZATION = 3
    MAX_POOL = 4
    SPLAT_LINEAR_INTERPOLATION = 5


class SparseTensorOperationMode(Enum):
    SEPARATE_COORDINATE_MANAGER = 0
    SHARE_COORDINATE_MANAGER = 1


class CoordinateMapType(Enum):
    CPU = 0
    CUDA = 1


class MinkowskiAlgorithm(Enum):
    DEFAULT = 0
    MEMORY_EFFICIENT = 1
    SPEED_OPTIMIZED = 2


class GPUMemoryAllocatorType(Enum):
    PYTORCH = 0
    CUDA = 1


class RegionType(Enum):
    HYPER_CUBE = 0
    HYPER_CROSS = 1
    CUSTOM = 2


class ConvolutionMode(Enum):
    DEFAULT = 0
    DIRECT_GEMM = 1
    COPY_GEMM = 2


class PoolingMode(Enum):
    LOCAL_SUM_POOLING = 0
    LOCAL_AVG_POOLING = 1
    LOCAL_MAX_POOLING = 2
    GLOBAL_SUM_POOLING_DEFAULT = 3
    GLOBAL_AVG_POOLING_DEFAULT = 4
    GLOBAL_MAX_POOLING_DEFAULT = 5
    GLOBAL_SUM_POOLING_KERNEL = 6
    GLOBAL_AVG_POOLING_KERNEL = 7
    GLOBAL_MAX_POOLING_KERNEL = 8
    GLOBAL_SUM_POOLING_PYTORCH_INDEX = 9
    GLOBAL_AVG_POOLING_PYTORCH_INDEX = 10
    GLOBAL_MAX_POOLING_PYTORCH_INDEX = 11


def _to_list(v, D: int, name: str) -> List[int]:
    if isinstance(v, torch.Tensor):
        v = v.tolist()
    if isinstance(v, (list, tuple)):
        if len(v) != D:
            raise ValueError(f"{name} must have {D} entries, got {v}")
        return [int(x) for x in v]
    return [int(v)] * D


# ---------------------------------------------------------------------------
# keys
# ---------------------------------------------------------------------------
class CoordinateMapKey:
    """(tensor_stride, string_id).  `CoordinateMapKey(coordinate_size)` builds an unset key that
    `set_key(stride, id)` completes (sparse_conv.py:399-401)."""

    def __init__(self, tensor_stride_or_size=None, string_id: str = ""):
        self._stride: Optional[Tuple[int, ...]] = None
        self._id: str = string_id
        self._coord_size: Optional[int] = None
        if isinstance(tensor_stride_or_size, int):
            self._coord_size = tensor_stride_or_size
        elif tensor_stride_or_size is not None:
            self._stride = tuple(int(s) for s in tensor_stride_or_size)
            self._coord_size = len(self._stride) + 1

    def set_key(self, tensor_stride, string_id: str = "") -> None:
        self._stride = tuple(int(s) for s in tensor_stride)
        self._id = string_id
        if self._coord_size is None:
            self._coord_size = len(self._stride) + 1

    def is_key_set(self) -> bool:
        return self._stride is not None

    def get_key(self):
        return list(self._stride) if self._stride is not None else None, self._id

    def get_tensor_stride(self) -> List[int]:
        if self._stride is None:
            raise RuntimeError("CoordinateMapKey: key is not set")
        return list(self._stride)

    def get_coordinate_size(self) -> int:
        return self._coord_size

    def _tuple(self):
        return (self._stride, self._id)

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and self._tuple() == other._tuple()

    def __hash__(self):
        return hash(self._tuple())

    def __repr__(self):
        return f"coordinate map key:{list(self._stride) if self._stride else None}" + (f":{self._id}" if self._id else "")


class KernelGenerator:
    """Kernel shape description (sparse_conv.py:438-452 reads these attributes)."""

    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose: bool = False,
                 region_type: RegionType = RegionType.HYPER_CUBE, region_offsets=None,
                 expand_coordinates: bool = False, axis_types=None, dimension: int = -1):
        assert dimension > 0, f"Invalid dimension: {dimension}"
        self.dimension = dimension
        self.kernel_size = _to_list(kernel_size, dimension, "kernel_size")
        self.kernel_stride = _to_list(stride, dimension, "stride")
        self.kernel_dilation = _to_list(dilation, dimension, "dilation")
        self.region_type = region_type
        self.region_offsets = region_offsets
        self.axis_types = axis_types
        self.is_transpose = is_transpose
        self.expand_coordinates = expand_coordinates
        if region_type != RegionType.HYPER_CUBE:
            raise NotImplementedError("only RegionType.HYPER_CUBE is built (the reference uses nothing else)")
        vol = 1
        for k in self.kernel_size:
            if k <= 0:
                raise ValueError(f"invalid kernel size {self.kernel_size}")
            vol *= k
        self.kernel_volume = vol
        # ME's (oddly named) flag: True when every stride is 1 (sparse_conv.py:322-335)
        self.requires_strided_coordinates = all(s == 1 for s in self.kernel_stride)

    def cache_key(self):
        return (tuple(self.kernel_size), tuple(self.kernel_stride), tuple(self.kernel_dilation), self.region_type)

    def __repr__(self):
        return (f"KernelGenerator(kernel_size={self.kernel_size}, stride={self.kernel_stride}, "
                f"dilation={self.kernel_dilation})")


# ---------------------------------------------------------------------------
# coordinate manager
# ---------------------------------------------------------------------------
class _FieldMap:
    __slots__ = ("coords", "size")

    def __init__(self, coords):
        self.coords = coords
        self.size = coords.shape[0]


class CoordinateManager:
    """Registry of coordinate maps and cached kernel maps of one forward pass."""

    default_prefetch_depth = 4  # stride-2 maps built ahead per stride() miss (1 = only the requested map)

    def __init__(self, D: int = 3, coordinate_map_type=None, allocator_type=None, minkowski_algorithm=None,
                 num_threads: int = -1):
        if D != 3:
            raise NotImplementedError("this engine is built for D=3 (batch,x,y,z) coordinates")
        self.D = D
        self.coordinate_map_type = coordinate_map_type or CoordinateMapType.CUDA
        self.minkowski_algorithm = minkowski_algorithm or MinkowskiAlgorithm.DEFAULT
        self._maps: Dict[CoordinateMapKey, ops.CoordMap] = {}
        self._fields: Dict[CoordinateMapKey, _FieldMap] = {}
        self._field_to_sparse: Dict[Tuple[CoordinateMapKey, CoordinateMapKey], torch.Tensor] = {}
        self._insert_aux: Dict[CoordinateMapKey, tuple] = {}   # key -> (first, inverse, count) of its creation
        self._stride_parent: Dict[CoordinateMapKey, tuple] = {}  # out key -> (in key, parent->child rows, count)
        self._kernel_maps: Dict[tuple, ops.KernelMap] = {}
        self._identity_maps: Dict[CoordinateMapKey, ops.KernelMap] = {}
        self._interp_maps: Dict[tuple, tuple] = {}             # (field key, map key) -> (rows [8,N], weights [8,N])
        self._field_batch: Dict[CoordinateMapKey, torch.Tensor] = {}
        self._row_perm: Dict[CoordinateMapKey, torch.Tensor] = {}   # maps whose rows were re-ordered: new row -> creation row
        self._n_batch: Optional[int] = None
        self._field_counter = 0
        self.prefetch_depth = CoordinateManager.default_prefetch_depth

    # ---- creation -----------------------------------------------------------
    def _unique_key(self, ts, string_id, table) -> CoordinateMapKey:
        key = CoordinateMapKey(ts, string_id)
        n = 0
        while key in table:
            n += 1
            key = CoordinateMapKey(ts, f"{string_id}-{n}" if string_id else f"{n}")
        return key

    def _maybe_reorder(self, key, cmap, first, inverse, count, parent_perm=None, parent_pos=None):
        """Engine-side row order (ops.reorder_rows_by_mask): called once, right after a map has been created and before
        anything refers to its rows.  `first` / `inverse` / `count` are the map's creation aux arrays; for a strided
        map they refer to the PARENT's rows, which may themselves just have been re-numbered (`parent_perm`,
        `parent_pos`).  The 3^3 stride-1 self map the decision needs is the one every convolution of the level uses:
        it is cached under its ME key, in the order the map ends up with.  Returns (first, inverse, count, perm, pos)."""
        if parent_perm is not None:
            inverse = inverse[parent_perm]                # indexed by parent rows: follow their new numbering
            first = parent_pos[first.long()]              # values are parent rows
        if not ops.sort_rows or cmap.size < ops.sort_min_rows:
            return first, inverse, count, None, None
        kg = KernelGenerator(kernel_size=3, stride=1, dilation=1, dimension=self.D)
        offs = ops.kernel_offsets(kg.kernel_size, cmap.tensor_stride, kg.kernel_dilation)
        km = ops.build_kernel_map(cmap, cmap, offs)
        res = ops.reorder_rows_by_mask(cmap, km)
        if res is not None:
            perm, pos = res
            first, count = first[perm], count[perm]
            safe = inverse.clamp_min(0).long()
            inverse = torch.where(inverse >= 0, pos[safe], inverse)   # (-1: a point outside the supported range)
            km = ops.build_kernel_map(cmap, cmap, offs)               # the same map in the new order
        self._kernel_maps[(key, key, kg.cache_key(), False, False)] = km
        return (first, inverse, count) + (res if res is not None else (None, None))

    def insert_and_map(self, coordinates: torch.Tensor, tensor_stride=1, string_id: str = ""):
        """int32 coordinates -> new map.  Returns (key, (unique_index, inverse_mapping)) (int64)."""
        ts = _to_list(tensor_stride, self.D, "tensor_stride")
        if coordinates.dtype != torch.int32:
            raise RuntimeError("insert_and_map expects int32 coordinates")
        cmap, first, inverse, count = ops.coords_insert(coordinates, L.SRC_INT, (1, 1, 1))
        cmap.tensor_stride = tuple(ts)
        key = self._unique_key(ts, string_id, self._maps)
        self._maps[key] = cmap
        self._insert_aux[key] = (first, inverse, count)
        return key, (first.long(), inverse.long())

    def insert_field(self, coordinates: torch.Tensor, tensor_stride=1, string_id: str = "") -> CoordinateMapKey:
        ts = _to_list(tensor_stride, self.D, "tensor_stride")
        if coordinates.dtype != torch.float32:
            coordinates = coordinates.float()
        self._field_counter += 1
        key = self._unique_key(ts, string_id or f"field{self._field_counter}", self._fields)
        self._fields[key] = _FieldMap(coordinates.contiguous())
        return key

    def field_to_sparse_insert_and_map(self, field_key: CoordinateMapKey, sparse_tensor_stride=1,
                                       sparse_string_id: str = ""):
        """Voxel quantisation of a coordinate field (TensorField.sparse, resnet.py:164)."""
        field = self._fields[field_key]
        ts = _to_list(sparse_tensor_stride, self.D, "tensor_stride")
        cmap, first, inverse, count = ops.coords_insert(field.coords, L.SRC_FLOAT, ts)
        key = self._unique_key(ts, sparse_string_id, self._maps)
        self._maps[key] = cmap
        first, inverse, count, perm, _ = self._maybe_reorder(key, cmap, first, inverse, count)
        if perm is not None:
            self._row_perm[key] = perm
        self._insert_aux[key] = (first, inverse, count)
        self._field_to_sparse[(field_key, key)] = inverse
        return key, (first, inverse, count)

    def exists_field_to_sparse(self, field_key, sparse_key) -> bool:
        return (field_key, sparse_key) in self._field_to_sparse

    def field_to_sparse_map(self, field_key, sparse_key) -> torch.Tensor:
        return self._field_to_sparse[(field_key, sparse_key)]

    def get_field_to_sparse_map(self, field_key, sparse_key):
        inv = self._field_to_sparse[(field_key, sparse_key)]
        return torch.arange(inv.shape[0], device=inv.device), inv.long()

    # ---- queries --------------------------------------------------------------
    def _map(self, key: CoordinateMapKey) -> ops.CoordMap:
        try:
            return self._maps[key]
        except KeyError:
            raise RuntimeError(f"coordinate map {key} does not exist in this manager") from None

    def exists_coordinate_map_key(self, key) -> bool:
        return key in self._maps

    def size(self, key: CoordinateMapKey) -> int:
        if key in self._maps:
            return self._maps[key].size
        if key in self._fields:
            return self._fields[key].size
        raise RuntimeError(f"coordinate map {key} does not exist in this manager")

    def get_coordinates(self, key: CoordinateMapKey) -> torch.Tensor:
        return self._map(key).coords

    def get_coordinate_field(self, key: CoordinateMapKey) -> torch.Tensor:
        return self._fields[key].coords

    def get_unique_coordinate_map_key(self, tensor_stride) -> CoordinateMapKey:
        ts = tuple(_to_list(tensor_stride, self.D, "tensor_stride"))
        found = [k for k in self._maps if k._stride == ts]
        if len(found) != 1:
            raise RuntimeError(f"{len(found)} coordinate maps with tensor stride {list(ts)}")
        return found[0]

    def number_of_unique_batch_indices(self) -> int:
        if self._n_batch is None:
            best = 0
            for cmap in self._maps.values():
                if cmap.size:
                    best = max(best, int(cmap.coords[:, 0].max().item()) + 1)  # host sync, once per manager
                    break
            if best == 0:                       # a pure point network (pointnet.py): only coordinate fields exist
                for field in self._fields.values():
                    if field.size:
                        best = max(best, int(torch.floor(field.coords[:, 0]).max().item()) + 1)
                        break
            self._n_batch = best
        return self._n_batch

    def field_batch_coordinates(self, field_key: CoordinateMapKey) -> torch.Tensor:
        """int32 [N,4] rows (floor(batch), 0, 0, 0) of a coordinate field: what the global pooling kernels read to
        find a point's instance (ME pools a TensorField per batch index too, pointnet.py:90,106)."""
        cached = self._field_batch.get(field_key)
        if cached is None:
            fc = self._fields[field_key].coords
            cached = torch.zeros((fc.shape[0], 4), dtype=torch.int32, device=fc.device)
            cached[:, 0] = torch.floor(fc[:, 0]).to(torch.int32)
            self._field_batch[field_key] = cached
        return cached

    # ---- stride ---------------------------------------------------------------
    def stride(self, in_key: CoordinateMapKey, stride, string_id: str = "") -> CoordinateMapKey:
        """Strided output map; an existing map with the resulting key is re-used (appendix A.3)."""
        s = _to_list(stride, self.D, "stride")
        if all(v == 1 for v in s):
            return in_key
        in_map = self._map(in_key)
        out_ts = [a * b for a, b in zip(in_map.tensor_stride, s)]
        out_key = CoordinateMapKey(out_ts, string_id)
        if out_key in self._maps:
            return out_key
        # A stride-2 request is (in every network of the reference) the first of a pyramid 2,4,8,16:
        # build `prefetch_depth` levels now with ONE host synchronisation, so that the host can enqueue
        # the whole encoder without stopping at every level to learn its row count.
        chain = [out_ts]
        if string_id == "" and all(v == 2 for v in s):
            while len(chain) < self.prefetch_depth:
                nxt = [2 * t for t in chain[-1]]
                if CoordinateMapKey(nxt, "") in self._maps or max(nxt) > 4096:
                    break
                chain.append(nxt)
        if len(chain) == 1:
            built = [ops.coords_insert(in_map.coords, L.SRC_STRIDE, out_ts)]
        else:
            built = ops.coords_insert_pyramid(in_map, chain)
        parent = in_key
        p_perm = p_pos = None     # (the levels were built from each other's creation order: re-number as we go down)
        for ts_l, (cmap, first, inverse, count) in zip(chain, built):
            key = CoordinateMapKey(ts_l, string_id)
            cmap.tensor_stride = tuple(ts_l)
            self._maps[key] = cmap
            first, inverse, count, p_perm, p_pos = self._maybe_reorder(key, cmap, first, inverse, count, p_perm, p_pos)
            if p_perm is not None:
                self._row_perm[key] = p_perm
            self._insert_aux[key] = (first, inverse, count)
            self._stride_parent[key] = (parent, inverse, count)
            parent = key
        return out_key

    def origin(self) -> CoordinateMapKey:
        """Origin map: one row (b,0,0,0) per batch index, ordered by batch index."""
        key = CoordinateMapKey([0] * self.D, "")
        if key not in self._maps:
            nb = self.number_of_unique_batch_indices()
            holders = list(self._maps.values()) or list(self._fields.values())
            dev = holders[0].coords.device
            coords = torch.zeros((nb, 4), dtype=torch.int32, device=dev)
            coords[:, 0] = torch.arange(nb, dtype=torch.int32, device=dev)
            cmap, _, _, _ = ops.coords_insert(coords, L.SRC_INT, (1, 1, 1))
            cmap.tensor_stride = (0,) * self.D
            self._maps[key] = cmap
        return key

    # ---- kernel maps ------------------------------------------------------------
    def get_kernel_map(self, in_key, out_key, kernel_generator: KernelGenerator, is_transpose: bool = False,
                       is_pool: bool = False) -> ops.KernelMap:
        """Cached dense kernel map between two existing maps (keyed like ME, appendix A.3)."""
        ck = (in_key, out_key, kernel_generator.cache_key(), bool(is_transpose), bool(is_pool))
        km = self._kernel_maps.get(ck)
        if km is not None:
            return km
        if is_transpose:
            # build the ordinary fine -> coarse map (fine = out_key here) and swap roles
            fwd = self.get_kernel_map(out_key, in_key, kernel_generator, False, is_pool)
            km = fwd.swapped()
        else:
            in_map, out_map = self._map(in_key), self._map(out_key)
            offs = ops.kernel_offsets(kernel_generator.kernel_size, in_map.tensor_stride,
                                      kernel_generator.kernel_dilation)
            km = ops.build_kernel_map(in_map, out_map, offs)
        self._kernel_maps[ck] = km
        return km

    def identity_map(self, key) -> ops.KernelMap:
        km = self._identity_maps.get(key)
        if km is None:
            cmap = self._map(key)
            nbr = torch.arange(cmap.size, dtype=torch.int32, device=cmap.coords.device).view(1, -1)
            km = ops.KernelMap(nbr, None, 1, cmap.size, cmap.size)
            km._nbr_t = nbr
            self._identity_maps[key] = km
        return km

    def kernel_map(self, in_key, out_key, stride=1, kernel_size=3, dilation=1, region_type=RegionType.HYPER_CUBE,
                   region_offset=None, is_transpose: bool = False, is_pool: bool = False):
        """ME API: dict {k: IntTensor[2, n_k]} (sparse_conv.py:90-96, :197-204)."""
        kg = KernelGenerator(kernel_size=kernel_size, stride=stride, dilation=dilation, region_type=region_type,
                             dimension=self.D)
        return self.get_kernel_map(in_key, out_key, kg, is_transpose, is_pool).pairs()

    def __repr__(self):
        lines = [f"CoordinateManager(D={self.D})"]
        for k, m in self._maps.items():
            lines.append(f"\t{k}:\tsize={m.size}")
        return "\n".join(lines)


# ---------------------------------------------------------------------------
# tensors
# ---------------------------------------------------------------------------
def _resolve_device(t: torch.Tensor, device) -> torch.device:
    if device is not None:
        return torch.device(device)
    return t.device


def _require_cuda(dev: torch.device, what: str):
    if dev.type != "cuda":
        raise RuntimeError(
            f"{what}: tensors are on '{dev}'. This engine has no CPU backend (sm_100a kernels only); "
            "move coordinates/features to a CUDA device or pass device='cuda'.")


class _Pending:
    """Deferred feature rows of a SparseTensor: [convolution] -> [BatchNorm [+ residual] [+ ReLU]].

    `conv` = (input rows, weight, bias, kernel map, precision, weight Parameter, offset bits); `x` = the input rows of
    a BatchNorm without a pending convolution; `bn` = the MinkowskiBatchNorm module."""
    __slots__ = ("conv", "x", "bn", "res", "relu", "gone", "grad", "bn_training", "cat")

    def __init__(self, conv=None, x=None, bn=None, cat=None):
        self.conv, self.x, self.bn = conv, x, bn
        self.cat = cat      # ME.cat: the parts' rows (possibly hollow)
        self.res = None
        self.relu = False
        self.gone = None
        # the modules were CALLED in this state; the rows are produced in it, whenever that happens
        self.grad = torch.is_grad_enabled()
        self.bn_training = bn.bn.training if bn is not None else None

    def run(self, want_fp32: bool = True) -> torch.Tensor:
        if self.gone is not None:
            raise RuntimeError(self.gone)
        with torch.set_grad_enabled(self.grad):
            return self._run(want_fp32)

    @staticmethod
    def _alias_wanted(src, x) -> bool:
        """hand the convolution's input back as an alias (ops "Residual gradients") when its rows take a gradient and
        the SparseTensor that holds them is known"""
        return bool(ops.fuse_residual_grad and src is not None and src._F is x and x.requires_grad
                    and torch.is_grad_enabled() and x.is_contiguous())

    def _run(self, want_fp32: bool) -> torch.Tensor:
        # rows that took a residual are a block output: the next block reads them as ITS residual in fp32
        want_fp32 = want_fp32 or self.res is not None
        if self.cat is not None:
            if want_fp32 or not ops.hollow_rows:
                return torch.cat([ops.ensure_filled(p) for p in self.cat], dim=1)
            return ops.CatFn.apply(*self.cat)
        if self.conv is not None:
            x, w, bias, km, precision, w_param, bits, src = self.conv
            alias = self._alias_wanted(src, x)
            if self.bn is not None and ops.conv_bn_fusable(x, w, bias, km, precision, bits):
                out = self.bn._run_fused(x, w, km, w_param, self.relu, self.res, want_fp32, self.bn_training, alias)
                feats = None
            else:
                out = ops.SparseConvFn.apply(x, w, bias, km, precision, w_param, bits, alias)
                feats = out[0] if alias else out
            if alias:
                src._F = out[1]     # later readers of the input rows (the residual branch) go through the alias
                out = out[0]
            if self.bn is None or feats is None:
                return out
        else:
            feats = self.x
        return self.bn._run(feats, self.relu, self.res, want_fp32, self.bn_training)


class Tensor:
    """Common base of SparseTensor and TensorField."""

    _F: torch.Tensor
    _manager: CoordinateManager
    quantization_mode: SparseTensorQuantizationMode

    # Lazily evaluated features.  A convolution or BatchNorm module returns a tensor whose rows are still PENDING
    # (`_Pending`: conv [+ BatchNorm [+ residual] [+ ReLU]]); the chain is resolved by its first reader, so that
    #   - a BatchNorm's `+= residual` and / or ReLU run in its apply kernel (one read + one write instead of three),
    #   - convolution + BatchNorm run as one autograd node (`ops.ConvBNFn`) in bf16 mode,
    #   - a reader that is itself a bf16 convolution asks for the bf16 operand copy only (`want_fp32=False`).
    _pending = None

    def _materialize(self, want_fp32: bool = True) -> None:
        pend = self._pending
        self._pending = None
        self._F = pend.run(want_fp32)

    @property
    def F(self) -> torch.Tensor:
        if self._F is None and self._pending is not None:
            self._materialize()
        return ops.ensure_filled(self._F) if self._F is not None else None

    def _operand(self) -> torch.Tensor:
        """Feature rows for a consumer that may read their bf16 operand copy instead (a convolution): pending rows
        are produced without their fp32 image when the precision mode allows, hollow rows are left hollow."""
        if self._F is None and self._pending is not None:
            self._materialize(want_fp32=False)
        return self._F

    @property
    def features(self) -> torch.Tensor:
        return self.F

    @property
    def feats(self) -> torch.Tensor:
        return self.F

    @property
    def coordinate_manager(self) -> CoordinateManager:
        return self._manager

    @property
    def D(self) -> int:
        return self._manager.D

    @property
    def dimension(self) -> int:
        return self._manager.D

    @property
    def device(self):
        return self.F.device

    @property
    def dtype(self):
        return self.F.dtype

    @property
    def shape(self):
        return self.F.shape

    @property
    def requires_grad(self):
        return self.F.requires_grad

    def requires_grad_(self, requires_grad: bool = True):
        self.F.requires_grad_(requires_grad)
        return self

    def size(self, *args):
        return self.F.size(*args)

    def __len__(self):
        return self.F.shape[0]


class SparseTensor(Tensor):
    def __init__(self, features: torch.Tensor, coordinates: Optional[torch.Tensor] = None, tensor_stride=1,
                 coordinate_map_key: Optional[CoordinateMapKey] = None,
                 coordinate_manager: Optional[CoordinateManager] = None,
                 quantization_mode: SparseTensorQuantizationMode = SparseTensorQuantizationMode.RANDOM_SUBSAMPLE,
                 allocator_type=None, minkowski_algorithm=None, requires_grad=None, device=None):
        assert isinstance(features, torch.Tensor), "Features must be a torch.Tensor"
        assert features.ndim == 2, f"features must be [N, C], got {tuple(features.shape)}"
        self.quantization_mode = quantization_mode
        self.unique_index = None
        self.inverse_mapping = None
        if coordinates is not None:
            assert coordinate_map_key is None, "give either coordinates or coordinate_map_key"
            dev = _resolve_device(features, device)
            _require_cuda(dev, "SparseTensor")
            if coordinates.shape[0] != features.shape[0]:
                raise RuntimeError("number of coordinates and features differ")
            coordinates = coordinates.to(dev)
            features = features.to(dev)
            if coordinates.dtype != torch.int32:
                if coordinates.is_floating_point():
                    warnings.warn("coordinates implicitly converted to torch.IntTensor (floor).")
                    coordinates = torch.floor(coordinates)
                coordinates = coordinates.int()
            if coordinate_manager is None:
                coordinate_manager = CoordinateManager(D=coordinates.shape[1] - 1)
            self._manager = coordinate_manager
            key, (first, inverse, count) = self._insert(coordinates, tensor_stride)
            self.coordinate_map_key = key
            m = coordinate_manager.size(key)
            mode = quantization_mode
            if mode == SparseTensorQuantizationMode.NO_QUANTIZATION or m == features.shape[0]:
                if m != features.shape[0]:
                    raise RuntimeError("NO_QUANTIZATION requires unique coordinates")
                self._F = features.float().contiguous() if features.dtype != torch.float32 else features.contiguous()
            else:
                code = {SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE: 0,
                        SparseTensorQuantizationMode.UNWEIGHTED_SUM: 1,
                        SparseTensorQuantizationMode.RANDOM_SUBSAMPLE: 2}.get(mode)
                if code is None:
                    raise NotImplementedError(f"quantization mode {mode} is not built")
                self._F = ops.SegmentReduceFn.apply(features.float(), inverse, count, first, m, code)
            self.unique_index = first
            self.inverse_mapping = inverse
        else:
            assert coordinate_map_key is not None and coordinate_manager is not None, \
                "coordinate_map_key and coordinate_manager are required when coordinates are not given"
            self._manager = coordinate_manager
            self.coordinate_map_key = coordinate_map_key
            if coordinate_map_key.is_key_set() and coordinate_manager.exists_coordinate_map_key(coordinate_map_key):
                if coordinate_manager.size(coordinate_map_key) != features.shape[0]:
                    raise RuntimeError(
                        f"features have {features.shape[0]} rows, coordinate map {coordinate_map_key} has "
                        f"{coordinate_manager.size(coordinate_map_key)}")
            self._F = features
        if requires_grad is not None:
            self._F.requires_grad_(requires_grad)

    @classmethod
    def _deferred(cls, pending: _Pending, coordinate_map_key, coordinate_manager) -> "SparseTensor":
        """A tensor whose feature rows are produced on first use by `pending.run()`."""
        t = object.__new__(cls)
        t.quantization_mode = SparseTensorQuantizationMode.RANDOM_SUBSAMPLE
        t.unique_index = None
        t.inverse_mapping = None
        t._manager = coordinate_manager
        t.coordinate_map_key = coordinate_map_key
        t._F = None
        t._pending = pending
        return t

    def _insert(self, coordinates, tensor_stride):
        mgr = self._manager
        ts = _to_list(tensor_stride, mgr.D, "tensor_stride")
        cmap, first, inverse, count = ops.coords_insert(coordinates.contiguous(), L.SRC_INT, (1, 1, 1))
        cmap.tensor_stride = tuple(ts)
        key = mgr._unique_key(ts, "", mgr._maps)
        mgr._maps[key] = cmap
        mgr._insert_aux[key] = (first, inverse, count)
        return key, (first, inverse, count)

    # ---- ME surface -------------------------------------------------------------
    @property
    def C(self) -> torch.Tensor:
        return self._manager.get_coordinates(self.coordinate_map_key)

    @property
    def coordinates(self) -> torch.Tensor:
        return self.C

    @property
    def tensor_stride(self) -> List[int]:
        return self.coordinate_map_key.get_tensor_stride()

    def _same_map(self, other: "SparseTensor"):
        if self._manager is not other._manager:
            raise AssertionError("tensors belong to different coordinate managers")
        if self.coordinate_map_key != other.coordinate_map_key:
            raise AssertionError(
                f"coordinate map keys differ: {self.coordinate_map_key} vs {other.coordinate_map_key}")

    def __iadd__(self, other):
        if isinstance(other, SparseTensor):
            self._same_map(other)
            pend = self._pending if self._F is None else None
            if pend is not None and pend.bn is not None and pend.res is None and not pend.relu and pend.gone is None:
                pend.res = other.F  # folded into the deferred BatchNorm apply (resnet_block.py:66)
            else:
                self._F = ops.AddFn.apply(self.F, other.F)
        else:
            self._F = self.F + other
        return self

    def __add__(self, other):
        if isinstance(other, SparseTensor):
            self._same_map(other)
            return SparseTensor(ops.AddFn.apply(self.F, other.F), coordinate_map_key=self.coordinate_map_key,
                                coordinate_manager=self._manager)
        return SparseTensor(self.F + other, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self._manager)

    def __mul__(self, other):
        if isinstance(other, SparseTensor):
            self._same_map(other)
            other = other.F
        return SparseTensor(self.F * other, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self._manager)

    def slice(self, X: "TensorField") -> "TensorField":
        """Features of this sparse tensor at the original points of `X` (res16unet.py:435)."""
        assert isinstance(X, TensorField), "slice expects a TensorField"
        if self.coordinate_map_key in getattr(X, "_identity_maps", ()):
            feats = self.F      # one point per voxel: the gather through the inverse map is the identity
        else:
            feats = ops.GatherRowsFn.apply(self.F, X.inverse_mapping(self.coordinate_map_key))
        return TensorField(feats,
                           coordinate_field_map_key=X.coordinate_field_map_key,
                           coordinate_manager=X.coordinate_manager, quantization_mode=X.quantization_mode)

    def interpolate(self, X: "TensorField") -> "TensorField":
        """Trilinear interpolation of this tensor's voxel features at the points of `X` (fcnn.py:200-203): each point
        reads the 8 voxels of THIS tensor's stride lattice around it; voxels that do not exist contribute zero."""
        assert isinstance(X, TensorField), "interpolate expects a TensorField"
        mgr = self._manager
        ck = (X.coordinate_field_map_key, self.coordinate_map_key)
        cached = mgr._interp_maps.get(ck)
        if cached is None:
            cached = ops.interp_map(mgr._map(self.coordinate_map_key), X.C)
            mgr._interp_maps[ck] = cached
        idx, w = cached
        return TensorField(ops.InterpolateFn.apply(self.F, idx, w),
                           coordinate_field_map_key=X.coordinate_field_map_key,
                           coordinate_manager=X.coordinate_manager, quantization_mode=X.quantization_mode)

    def features_at(self, batch_index: int) -> torch.Tensor:
        return self.F[self.C[:, 0] == batch_index]

    def coordinates_at(self, batch_index: int) -> torch.Tensor:
        C = self.C
        return C[C[:, 0] == batch_index, 1:]

    @property
    def decomposed_features(self):
        nb = self._manager.number_of_unique_batch_indices()
        return [self.features_at(b) for b in range(nb)]

    @property
    def decomposed_coordinates(self):
        nb = self._manager.number_of_unique_batch_indices()
        return [self.coordinates_at(b) for b in range(nb)]

    def detach(self):
        return SparseTensor(self.F.detach(), coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self._manager)

    def __repr__(self):
        return (f"SparseTensor(\n  coordinates={self.C}\n  features={self.F}\n  "
                f"coordinate_map_key={self.coordinate_map_key}\n)")


class TensorField(Tensor):
    def __init__(self, features: torch.Tensor, coordinates: Optional[torch.Tensor] = None, tensor_stride=1,
                 coordinate_field_map_key: Optional[CoordinateMapKey] = None,
                 coordinate_manager: Optional[CoordinateManager] = None,
                 quantization_mode: SparseTensorQuantizationMode = SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE,
                 allocator_type=None, minkowski_algorithm=None, requires_grad=None, device=None):
        assert isinstance(features, torch.Tensor), "Features must be a torch.Tensor"
        assert features.ndim == 2, f"features must be [N, C], got {tuple(features.shape)}"
        self.quantization_mode = quantization_mode
        self._inverse_mapping: Dict[CoordinateMapKey, torch.Tensor] = {}
        self._identity_maps = set()   # sparse maps with one point per voxel: row r of the map IS point r
        if coordinates is not None:
            dev = _resolve_device(features, device)
            _require_cuda(dev, "TensorField")
            if coordinates.shape[0] != features.shape[0]:
                raise RuntimeError("number of coordinates and features differ")
            coordinates = coordinates.to(dev)
            features = features.to(dev)
            if coordinate_manager is None:
                coordinate_manager = CoordinateManager(D=coordinates.shape[1] - 1)
            self._manager = coordinate_manager
            self.coordinate_field_map_key = coordinate_manager.insert_field(coordinates.float(), tensor_stride)
        else:
            assert coordinate_field_map_key is not None and coordinate_manager is not None, \
                "coordinate_field_map_key and coordinate_manager are required when coordinates are not given"
            self._manager = coordinate_manager
            self.coordinate_field_map_key = coordinate_field_map_key
        self._F = features
        if requires_grad is not None:
            self._F.requires_grad_(requires_grad)

    @property
    def C(self) -> torch.Tensor:
        return self._manager.get_coordinate_field(self.coordinate_field_map_key)

    @property
    def coordinates(self) -> torch.Tensor:
        return self.C

    def sparse(self, tensor_stride=1, coordinate_map_key: Optional[CoordinateMapKey] = None,
               quantization_mode: Optional[SparseTensorQuantizationMode] = None) -> SparseTensor:
        """Voxel quantisation + hashing + per-voxel feature reduction (resnet.py:164, res16unet.py:392)."""
        mode = quantization_mode if quantization_mode is not None else self.quantization_mode
        code = {SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE: 0,
                SparseTensorQuantizationMode.UNWEIGHTED_SUM: 1,
                SparseTensorQuantizationMode.RANDOM_SUBSAMPLE: 2}.get(mode)
        if code is None:
            raise NotImplementedError(f"quantization mode {mode} is not built")
        mgr = self._manager
        if coordinate_map_key is not None and mgr.exists_field_to_sparse(self.coordinate_field_map_key,
                                                                         coordinate_map_key):
            key = coordinate_map_key
            first, inverse, count = mgr._insert_aux[key]
        else:
            key, (first, inverse, count) = mgr.field_to_sparse_insert_and_map(
                self.coordinate_field_map_key, tensor_stride)
        self._inverse_mapping[key] = inverse
        m = mgr.size(key)
        feats = self._F if self._F.dtype == torch.float32 else self._F.float()
        perm = mgr._row_perm.get(key)
        if m == feats.shape[0] and inverse.shape[0] == m and perm is not None:
            # one point per voxel, rows re-ordered by the engine (ops.reorder_rows_by_mask): row j is point perm[j]
            F = ops.GatherRowsFn.apply(feats.contiguous(), perm.int() if perm.dtype != torch.int32 else perm)
        elif m == feats.shape[0] and inverse.shape[0] == m:
            # No two points share a voxel (what the reference's plenoxel loaders deliver: one record per occupied
            # cell).  Rows are in first-occurrence order, so the map's row r is point r, the inverse map is the
            # identity and every reduction of one value is that value: no pass over the features, here or in slice().
            self._identity_maps.add(key)
            F = feats.contiguous()
        else:
            F = ops.SegmentReduceFn.apply(feats, inverse, count, first, m, code)
        return SparseTensor(F, coordinate_map_key=key, coordinate_manager=mgr)

    def splat(self) -> SparseTensor:
        """Spread every point's features over the 8 voxels around it with trilinear weights (fcnn.py:186): the result
        lives on a new map holding all touched voxels (in order of first touch); `Y.interpolate(X)` is the way back."""
        mgr = self._manager
        field_ts = self.coordinate_field_map_key.get_tensor_stride()
        ts = [max(int(t), 1) for t in field_ts]
        q = self.C
        n = q.shape[0]
        lower, w = ops.interp_corners(q, ts)
        offs = torch.tensor(ops.corner_offsets(ts), dtype=torch.int32, device=q.device)          # [8, 3]
        corners = lower[:, None, :].repeat(1, 8, 1)
        corners[:, :, 1:] += offs[None]
        key, _ = mgr.insert_and_map(corners.view(-1, 4), tensor_stride=ts)
        idx = mgr._insert_aux[key][1].view(n, 8).t().contiguous()                                 # rows, [8, N]
        mgr._interp_maps[(self.coordinate_field_map_key, key)] = (idx, w)
        feats = self._F if self._F.dtype == torch.float32 else self._F.float()
        F = ops.SplatFn.apply(feats, idx, w, mgr.size(key))
        return SparseTensor(F, coordinate_map_key=key, coordinate_manager=mgr)

    def inverse_mapping(self, sparse_key: CoordinateMapKey) -> torch.Tensor:
        """Row of `sparse_key`'s map each point falls into.  For the map this field's `.sparse()` created it is the
        stored inverse map; for any other map (a strided level: `y2.slice(x)`, fcnn.py:162-165) every point's voxel
        `floor(c / ts) * ts` is looked up in that map's hash table."""
        inv = self._inverse_mapping.get(sparse_key)
        if inv is None:
            mgr = self._manager
            if mgr.exists_field_to_sparse(self.coordinate_field_map_key, sparse_key):
                inv = mgr.field_to_sparse_map(self.coordinate_field_map_key, sparse_key)
            else:
                cmap = mgr._map(sparse_key)
                lower, _ = ops.interp_corners(self.C, cmap.tensor_stride)
                km = ops.build_kernel_map(cmap, ops.CoordMap(lower, None, 0, lower.shape[0], cmap.tensor_stride),
                                          [(0, 0, 0)])
                inv = km.nbr[0]
                if int(km.tap_count[0].item()) != lower.shape[0]:          # host sync, once per (field, map)
                    raise RuntimeError(f"slice(): some points of the field have no voxel in {sparse_key}")
                mgr._field_to_sparse[(self.coordinate_field_map_key, sparse_key)] = inv
            self._inverse_mapping[sparse_key] = inv
        return inv

    def __repr__(self):
        return f"TensorField(\n  coordinates={self.C}\n  features={self._F}\n)"
