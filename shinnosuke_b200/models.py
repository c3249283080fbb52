"""Sequential / Model training drivers on the device.

API of the reference's models.py (Sequential :12-266, Model :272-588): compile / fit / predict /
evaluate / calc_gradients / save / load, `train_loss` / `train_acc` histories.  What changes:

  * compile() takes `precision=` ('fp32' CUDA-core tier, 'tf32' / 'bf16' tensor-core tiers);
  * the first fit/predict places every trainable Variable in ONE flat fp32 parameter arena with
    a twin gradient arena, so the optimizer is one fused kernel and — under
    ``torch.distributed`` (one process per GPU) — the gradient exchange is one NCCL all-reduce;
  * fit() shards each minibatch across the ranks (contiguous rows); the loss
    gradient is divided by the GLOBAL row count so the summed shard gradients are exactly the
    reference's full-batch gradients;
  * the whole step (forward sweep, loss, backward sweep, all-reduce, update) is captured once
    per batch shape in a CUDA graph and replayed; the inputs are copied host->device from pinned
    memory every step and the (loss, acc) pair is copied back every step without blocking.
"""
import os
import pickle
import time

import numpy as np

from .device import Arena, DeviceTensor, as_device, current_stream, empty, require_cuda, torch, zeros
from ._lib import lib, PRECISION
from .layers.Base import Layer, Variable, Input, flush_deferred
from .utils.MiniBatch import batch_permutation, batch_slices
from .utils.Objectives import get_objective
from .utils.Optimizers import get_optimizer, Momentum
from .comm import Parallel


class VariableList(list):
    """`trainable_variables` with a handle on the arena that backs them."""
    arena = None


def _dist():
    t = torch()
    if t.distributed.is_available() and t.distributed.is_initialized():
        return t.distributed
    return None


class _StepGraph(object):
    """An instantiated CUDA graph of one training step (sn_graph_* in the C ABI)."""

    def __init__(self, handle):
        self.handle = handle

    def replay(self):
        lib.sn_graph_launch(self.handle, current_stream())

    def __del__(self):
        try:
            if self.handle is not None:
                lib.sn_graph_destroy(self.handle)
        except Exception:
            pass


class _Trainer(object):
    """Everything Sequential and Model share once the layer order is known."""

    process_bar_nums = 30
    process_bar_trained = '='
    process_bar_untrain = '*'

    def _init_histories(self):
        self.train_loss = []
        self.train_acc = []
        self.valid_loss = []
        self.valid_acc = []
        self.precision = 'fp32'
        self.use_cuda_graph = True
        self._arena = None
        self._graphs = {}
        self._warm = set()
        self._metrics_out = None
        self._fit_res = None
        self._split = None
        self._comm_stream = None
        self._par = None            # data-parallel context (comm.Parallel), created with the arena
        self._replicas = None       # fit(n_gpus=G): this model and its G - 1 copies on the other devices
        self._cap_stream = None
        self._bf16_shadow = None
        self._prep_tasks = {}
        self._pinned_user = {}      # id -> (array, ptr): user arrays page-locked in place by fit(shuffle=False)

    # ---- to be provided by the subclass ---------------------------------
    def _forward_order(self):
        raise NotImplementedError

    def _backward_order(self):
        raise NotImplementedError

    def _input_layer(self):
        raise NotImplementedError

    def _output_layer(self):
        raise NotImplementedError

    # ---- compile ---------------------------------------------------------
    def _finish_compile(self, optimizer, loss, precision, fuse=True, sync_bn=None):
        if precision not in PRECISION:
            raise ValueError('precision must be one of %s' % sorted(PRECISION))
        self.precision = precision
        from .layers.Normalization import BatchNormalization
        from .layers.Convolution import MaxPooling2D, Conv2D
        for layer in self._forward_order():
            layer.precision = precision
            if isinstance(layer, BatchNormalization):
                # cross-replica statistics: the reference normalises over the WHOLE
                # minibatch (layers/Normalization.py:116-128), so exact data parallelism is the default
                # whenever there is more than one replica; sync_bn=False keeps per-replica statistics
                layer._sync_bn = True if sync_bn is None else bool(sync_bn)
                # BatchNormalization -> MaxPooling2D((2,2)) with no other consumer runs as one fused pass
                nxt = layer.outbound_layers[0] if len(layer.outbound_layers) == 1 else None
                ok = (fuse and isinstance(nxt, MaxPooling2D) and tuple(nxt.pool_size) == (2, 2) and nxt.stride == 2
                      and len(nxt.inbounds) == 1)
                layer._fuse_with_pool = nxt if ok else None
                # Conv2D -> BatchNormalization with no other consumer: the tensor-core convolution
                # epilogues reduce the batch statistics, so the BN skips its own pass over the map
                prev = layer.inbounds[0] if len(layer.inbounds) == 1 else None
                layer._stats_from = None
                if isinstance(prev, Conv2D):
                    feeds = fuse and precision in ('tf32', 'bf16') and len(prev.outbound_layers) == 1
                    prev._stats_to = layer if feeds else None
        # BatchNormalization / Add -> Activation('relu') with no other consumer: the ReLU rides on the
        # producing kernel's output pass and the Activation layer only hands the buffer on
        from .layers.Activation import Activation
        from .layers.Base import Add
        for layer in self._forward_order():
            if isinstance(layer, (BatchNormalization, Add)):
                layer._relu_out = None
        for layer in self._forward_order():
            if isinstance(layer, Activation):
                layer._folded_into = None
                prev = layer.inbounds[0] if len(layer.inbounds) == 1 else None
                ok = (fuse and layer._kind() == 'relu' and isinstance(prev, (BatchNormalization, Add))
                      and len(prev.outbound_layers) == 1 and getattr(prev, '_fuse_with_pool', None) is None)
                if ok:
                    prev._relu_out = layer
                    layer._folded_into = prev
        # Flatten -> Dense with nobody else on either side: the NHWC map is handed through as it is and
        # the Dense keeps its weight rows in (i,j,c) order (no transpose kernel forward or backward)
        from .layers.FC import Flatten, Dense
        for layer in self._forward_order():
            if isinstance(layer, Flatten):
                nxt = layer.outbound_layers[0] if len(layer.outbound_layers) == 1 else None
                shp = layer.inbounds[0].output_shape if len(layer.inbounds) == 1 else None
                ok = (fuse and isinstance(nxt, Dense) and len(nxt.inbounds) == 1 and shp is not None and len(shp) == 4
                      and shp[2] * shp[3] > 1 and shp[1] > 1)
                layer._hwc_passthrough = bool(ok)
                if isinstance(nxt, Dense) and len(nxt.inbounds) == 1 and shp is not None and len(shp) == 4:
                    nxt.variables[0].relayout(('dense_w_hwc',) + tuple(int(v) for v in shp[1:]) if ok else 'dense_w')
        self.loss = get_objective(loss)
        self.optimizer = get_optimizer(optimizer)
        # a re-compile (another optimizer, precision, fusion plan) invalidates every captured step
        self._graphs = {}
        self._warm = set()

    def _materialize(self):
        """Place all trainable variables in the arena (once)."""
        if self._arena is not None:
            return
        require_cuda()
        tv = [v for v in self.trainable_variables if isinstance(v, Variable)]
        sizes = [v.size for v in tv]
        _, _, total = Arena.layout(sizes, Arena.ALIGN)
        if self._par is None:
            self._par = Parallel.from_torch_distributed(total)
        par = self._par
        keeps_grads = isinstance(self.optimizer, Momentum) or getattr(self.optimizer, 'needs_gstep', False)
        # the buffer that is summed over the replicas every step lives in the communicator's peer-mapped
        # window (all-reduced in place, zero-copy): the gradient arena itself, or `gstep` for the optimizers
        # that never zero the gradients (each step's all-reduced gradients are folded into arena.g by the
        # update kernel)
        in_window = par.world > 1 and par.own_kernels
        arena = Arena(sizes, extra=Arena.ALIGN, g_alloc=par.grad.carve_f32 if (in_window and not keeps_grads) else None)
        for i, v in enumerate(tv):
            v.bind(arena.slice('w', i, v.size), arena.slice('g', i, v.size))
        self._arena = arena
        vl = VariableList(tv)
        vl.arena = arena
        self.trainable_variables = vl
        self._metrics_out = zeros(4)          # two (loss, acc) slots: fit() reads one back while the next step writes the other
        from .layers.Convolution import Conv2D
        for layer in self._forward_order():
            layer._par = par
            # data parallel: a convolution postpones its weight gradient into the upstream BatchNormalization's exchange
            layer._defer_wgrad = bool(par.world > 1 and par.own_kernels and isinstance(layer, Conv2D))
        if par.world > 1:
            if keeps_grads:
                arena.gstep = par.grad.carve_f32(arena.total) if in_window else zeros(arena.total)
            self._sync_initial_weights()
        lib.sn_stream_sync(current_stream())

    def _sync_initial_weights(self):
        """Replicas must start from identical weights: rank 0's initial values win."""
        par = self._par
        if par.dist is not None:
            par.dist.broadcast(self._arena.w, src=0)

    # ---- forward / backward sweeps (models.py:124-129, 82-83, 445-449) -----
    def _forward_device(self, x_dev, is_training):
        self._input_layer().input_tensor = x_dev
        for layer in self._forward_order():
            layer.forward(is_training=is_training)
        return self._output_layer().output_tensor

    def predict(self, X, is_training=True):
        self._materialize()
        return self._forward_device(self._to_device_input(X), is_training)

    def _to_device_input(self, X):
        if isinstance(X, DeviceTensor):
            return X
        return as_device(np.asarray(X))

    def calc_gradients(self, y_hat, y_true, denom_rows=0, metrics_ptr=None):
        out = self._output_layer()
        out._seed_grads(self.loss.backward(y_hat, as_device(y_true), metrics_ptr, denom_rows))
        for layer in self._backward_order():
            layer.backward()
        flush_deferred()

    # ---- one training step on device-resident shards ---------------------
    def _grad_arena_for_step(self):
        a = self._arena
        return a.gstep if a.gstep is not None else a.g

    def _prepare_bf16_weights(self, n_batch):
        """bf16 tier: ONE launch casts the whole parameter arena to its bf16 shadow (every Conv2D's fprop
        filters are slices of it) and writes the flipped + transposed dgrad operand of every
        convolution whose dgrad runs on bf16 operands at this batch size (three casts + three flips
        for config 3 otherwise).  Returns the layers it served; they are released at the end of the step."""
        from .layers.Convolution import Conv2D
        a = self._arena
        t = torch()
        if getattr(self, '_bf16_shadow', None) is None:
            self._bf16_shadow = empty(a.total, t.bfloat16)
            self._prep_tasks = {}
        index = {id(v): i for i, v in enumerate(self.trainable_variables)}
        convs = [l for l in self._forward_order() if isinstance(l, Conv2D) and l.variables and id(l.variables[0]) in index
                 and l.input_shape is not None and len(l.input_shape) == 4]

        def flips_dgrad(l):             # Conv2D._bf16_grad_plan's rule for the dgrad operand, from static shapes
            _, C, H, Wd = l.input_shape
            kh, kw = l.filter_size
            geom = (n_batch, C, H, Wd, l.filter_nums, kh, kw, l.stride, l.pad_size[0], l.pad_size[1])
            return any(p.require_grads for p in l.inbounds) and l._bf16_ok(1, geom)
        key = n_batch
        if key not in self._prep_tasks:
            flips = [l for l in convs if flips_dgrad(l)]
            rows, start = [], 0
            for l in flips:
                W = l.variables[0]
                F, C = l.filter_nums, W.output_shape[1]
                kh, kw = l.filter_size
                wt = l._bf16_buf('wt_bf16', W.size)
                units = kh * kw * ((F + 31) // 32) * ((C + 31) // 32)       # one 32x32 (f, c) tile of one tap each
                rows.append([a.offsets[index[id(W)]], wt.data_ptr(), F, C, kh, kw, start, units])
                start += units
            table = t.tensor(rows if rows else [[0] * 8], dtype=t.int64, device='cuda')
            self._prep_tasks[key] = (table, len(rows), start, set(id(l) for l in flips))
        table, ntasks, flip_elems, flipped = self._prep_tasks[key]
        lib.sn_conv_weights_prepare_bf16(a.w.data_ptr(), self._bf16_shadow.data_ptr(), a.param_floats, table.data_ptr(), ntasks,
                                         flip_elems, current_stream())
        for l in convs:
            W = l.variables[0]
            l._w_prepared = (self._bf16_shadow.data_ptr() + 2 * a.offsets[index[id(W)]], id(l) in flipped)
        return convs

    def _step_body(self, x_dev, y_dev, global_rows, mslot=0):
        """forward, loss, backward, gradient all-reduce, update.  No host sync inside: this is
        what gets captured in the CUDA graph."""
        a = self._arena
        st = current_stream()
        dist = self._par
        world = dist.world
        garena = self._grad_arena_for_step()
        metrics = garena[a.param_floats:a.param_floats + 2]
        if a.gstep is not None:
            # data-parallel Momentum: layers accumulate this step's gradients into gstep
            for i, v in enumerate(self.trainable_variables):
                v._grads = DeviceTensor(a.slice('gstep', i, v.size), v.output_shape, v.layout or 'flat')
        if world > 1:
            for layer in self._forward_order():
                if getattr(layer, '_sync_bn', False):
                    layer._global_batch = global_rows
        prepared = self._prepare_bf16_weights(int(x_dev.shape[0])) if self.precision == 'bf16' else ()
        try:
            self._step_rest(x_dev, y_dev, global_rows, mslot, a, st, dist, world, garena, metrics)
        finally:
            for layer in prepared:
                layer._w_prepared = None

    def _step_rest(self, x_dev, y_dev, global_rows, mslot, a, st, dist, world, garena, metrics):
        y_hat = self._forward_device(x_dev, True)
        updated = False
        if world > 1:
            updated = self._backward_allreduce(y_hat, y_dev, global_rows, metrics, garena, dist)
        else:
            self.calc_gradients(y_hat, y_dev, denom_rows=global_rows, metrics_ptr=metrics.data_ptr())
        lib.sn_memcpy_d2d(self._metrics_out.data_ptr() + 8 * mslot, metrics.data_ptr(), 8, st)
        self.optimizer.grad_scale = 1.0                     # shards already divide by the global rows
        if updated:
            self.optimizer.iterations += self.optimizer.ITER_PER_UPDATE
        else:
            self.optimizer.update(self.trainable_variables)
        # the optimizer touches arena.g[:param_floats] only; clear the metric tail for the next step
        lib.sn_memset(metrics.data_ptr(), 0, 8, st)

    def _overlap_split(self):
        """Where the backward pass can start reducing: (index into the backward order after which the
        arena suffix [lo, end) is final, lo).  The arena is laid out in forward order, so the layers
        visited first in backward own its tail; the split is the first point where that tail holds
        >= 80 % of the parameters while at least one layer with variables is still to come."""
        if getattr(self, '_split', None) is not None:
            return self._split
        a = self._arena
        index = {id(v): i for i, v in enumerate(self.trainable_variables)}
        order = self._backward_order()
        ends = []                   # per layer: arena ranges of its trainable variables
        for layer in order:
            ends.append([(a.offsets[index[id(v)]], a.offsets[index[id(v)]] + v.size) for v in layer.variables if id(v) in index])
        split = (None, 0)
        for k in range(len(order) - 1):
            rest = [r for rs in ends[k + 1:] for r in rs]
            if not rest:
                break
            lo = max(e for _, e in rest)
            lo = (lo + a.ALIGN - 1) // a.ALIGN * a.ALIGN
            done = [r for rs in ends[:k + 1] for r in rs]
            if done and min(b for b, _ in done) >= lo and a.param_floats - lo >= 0.8 * a.param_floats:
                split = (k, lo)
                break
        self._split = split
        return split

    def _backward_allreduce(self, y_hat, y_dev, global_rows, metrics, garena, dist):
        """Backward pass of one data-parallel shard with the gradient all-reduce (NCCL sum over
        NVLink) overlapped: as soon as the tail of the arena is final (the deep layers, most of the
        parameters) it is reduced on a side stream, together with the loss/accuracy sums riding in
        the arena's last words, while the remaining layers run their backward; the short head of
        the arena follows at the end."""
        t = torch()
        a = self._arena
        par = dist
        k_split, lo = self._overlap_split()
        out = self._output_layer()
        out._seed_grads(self.loss.backward(y_hat, as_device(y_dev), metrics.data_ptr(), global_rows))
        main = t.cuda.current_stream()
        # Default: ONE exchange after the backward sweep.  Measured at 2 GPUs (config 3, 512 per GPU): 0.545 ms
        # per step against 0.552 with the tail of the arena reduced on a side stream under the remaining
        # backward kernels — the fork / join edges break the programmatic-dependent-launch chain of the
        # step and the exchange CTAs compete with the persistent convolution kernels, which costs more than
        # the ~20 us exchange hides.  SN_COMM_OVERLAP=1 (or the NCCL fallback, whose all-reduce is 3-4x
        # slower) keeps the overlapped form.
        overlap = os.environ.get('SN_COMM_OVERLAP', '0' if par.own_kernels else '1') == '1'
        if k_split is None or not overlap:
            for layer in self._backward_order():
                layer.backward()
            flush_deferred()
            from .utils.Optimizers import StochasticGradientDescent
            if par.own_kernels and type(self.optimizer) is StochasticGradientDescent and garena is a.g \
                    and os.environ.get('SN_FUSE_SGD', '1') == '1':
                # gradient exchange + SGD update in ONE kernel over peer memory (csrc/comm.cu)
                lib.sn_allreduce_sgd_update(par.grad.handle, garena.data_ptr(), a.w.data_ptr(), garena.numel(), a.param_floats,
                                            float(self.optimizer.lr), main.cuda_stream)
                return True
            par.allreduce_grads(garena, main.cuda_stream)
            return False
        if getattr(self, '_comm_stream', None) is None:
            self._comm_stream = t.cuda.Stream()
        comm = self._comm_stream
        for k, layer in enumerate(self._backward_order()):
            layer.backward()
            if k == k_split:
                flush_deferred()
                ready = t.cuda.Event()
                ready.record(main)
                comm.wait_event(ready)
                with t.cuda.stream(comm):
                    par.allreduce_grads(garena[lo:], comm.cuda_stream)
        flush_deferred()
        # the head follows on the SAME side stream: one window's flags serve one exchange at a time
        ready = t.cuda.Event()
        ready.record(main)
        comm.wait_event(ready)
        with t.cuda.stream(comm):
            par.allreduce_grads(garena[:lo], comm.cuda_stream)
        done = t.cuda.Event()
        done.record(comm)
        main.wait_event(done)
        return False

    def _run_step(self, x_dev, y_dev, global_rows, mslot=0):
        if not self.use_cuda_graph:
            self._step_body(x_dev, y_dev, global_rows, mslot)
            return
        key = (x_dev.ptr, y_dev.ptr, x_dev.shape, global_rows, mslot, self.optimizer.hyper_key())
        graph = self._graphs.get(key)
        if graph is None:
            if key not in self._warm:
                # first step of this shape runs eagerly: allocates every layer buffer and loads
                # every kernel, so that nothing but launches happens during capture
                self._warm.add(key)
                self._step_body(x_dev, y_dev, global_rows, mslot)
                return
            graph = self._capture(x_dev, y_dev, global_rows, mslot)
            self._graphs[key] = graph
            # capture ran update() once on the host (iterations += 1) without executing anything
            self.optimizer.iterations -= self.optimizer.ITER_PER_UPDATE
        graph.replay()
        self.optimizer.iterations += self.optimizer.ITER_PER_UPDATE

    def _capture(self, x_dev, y_dev, global_rows, mslot):
        """Capture one training step as a CUDA graph through the C ABI (sn_graph_begin / sn_graph_end) on
        a side stream; every exchange inside is one of our own kernels.  Only the NCCL fallback
        (SN_COMM=nccl) goes through torch.cuda.CUDAGraph, which knows how to capture NCCL."""
        import ctypes
        t = torch()
        par = self._par
        if par.world > 1 and not par.own_kernels:
            graph = t.cuda.CUDAGraph()
            with t.cuda.graph(graph):
                self._step_body(x_dev, y_dev, global_rows, mslot)
            return graph
        if self._cap_stream is None:
            self._cap_stream = t.cuda.Stream()
        cap, cur = self._cap_stream, t.cuda.current_stream()
        cap.wait_stream(cur)
        handle = ctypes.c_void_p()
        with t.cuda.stream(cap):
            lib.sn_graph_begin(cap.cuda_stream)
            try:
                self._step_body(x_dev, y_dev, global_rows, mslot)
            finally:
                lib.sn_graph_end(cap.cuda_stream, ctypes.byref(handle))
        cur.wait_stream(cap)
        return _StepGraph(handle)

    # ---- fit (models.py:47-119 / :370-431) --------------------------------
    def _pin_in_place(self, arr):
        """Page-lock a caller-owned float32 array (cached: locking costs ~0.3 ms/MB, a training step
        ~1 ms).  Returns False if the driver refuses (e.g. memory already locked by someone else)."""
        ptr = arr.ctypes.data
        hit = self._pinned_user.get(ptr)
        if hit is not None:
            if hit[1] >= arr.nbytes:
                return True
            try:                                              # same base, longer array: re-lock the longer range
                lib.sn_host_unregister(ptr)
            except Exception:
                pass
            del self._pinned_user[ptr]
        if len(self._pinned_user) >= 8:                       # bounded cache
            old_ptr, (_, _) = next(iter(self._pinned_user.items()))
            try:
                lib.sn_host_unregister(old_ptr)
            except Exception:
                pass
            del self._pinned_user[old_ptr]
        try:
            lib.sn_host_register(ptr, arr.nbytes)
        except Exception:
            return False
        self._pinned_user[ptr] = (arr, arr.nbytes)
        return True

    def _fit_resources(self, n_loc_max, x_row, y_row, n_batches):
        """Pinned staging ring, copy stream and metric buffers; cached on the model because
        page-locking host memory is far slower than a training step."""
        t = torch()
        key = (x_row, y_row)
        res = self._fit_res
        if res is None or res['key'] != key or res['cap'] < n_loc_max:
            depth = 4
            res = dict(key=key, cap=n_loc_max, depth=depth,
                       x=[t.empty((n_loc_max, x_row), dtype=t.float32).pin_memory() for _ in range(depth)],
                       y=[t.empty((n_loc_max, y_row), dtype=t.float32).pin_memory() for _ in range(depth)],
                       copy_stream=t.cuda.Stream(), metrics=None,
                       h2d_done=[t.cuda.Event() for _ in range(2)], compute_done=[t.cuda.Event() for _ in range(2)],
                       d2h_stream=t.cuda.Stream(), d2h_done=[t.cuda.Event() for _ in range(2)],
                       slot_copied=[t.cuda.Event() for _ in range(depth)])
            self._fit_res = res
        if res['metrics'] is None or res['metrics'].shape[0] < n_batches:
            res['metrics'] = t.zeros((n_batches, 2), dtype=t.float32).pin_memory()
        return res

    # ---- fit(n_gpus=G): one process, G replicas -----------------------------------------
    def _fit_replicated(self, n_gpus, X, Y, batch_size, epochs, shuffle, validation_data, validation_ratio, verbose):
        """models.py fit() sharding each minibatch over the GPUs of the box BY ITSELF: replica r lives on device r
        and is driven by its own host thread through the ordinary fit() loop (rank r's rows of every minibatch, its
        own copy stream, pinned staging and captured step graph); the replicas meet only inside the exchange kernels
        of csrc/comm.cu (peer access between the devices of this process, sn_comm_init_all).  Threads, not a
        sequential loop over devices: a replica's first step allocates buffers and loads kernels while another
        replica's exchange kernel is already spinning for it, and nothing on the host may wait in between."""
        import copy
        import threading
        t = require_cuda()
        if _dist() is not None and _dist().get_world_size() > 1:
            raise RuntimeError('fit(n_gpus=...) drives all GPUs from this process; under torchrun every rank already owns one')
        have = t.cuda.device_count()
        if n_gpus > have:
            raise ValueError('fit(n_gpus=%d): only %d CUDA devices are visible' % (n_gpus, have))
        if n_gpus > 8:
            raise ValueError('fit(n_gpus=...): at most 8 replicas (one NVLink / NVSwitch box)')
        reps = getattr(self, '_replicas', None)
        if reps is None or len(reps) != n_gpus:
            # (re)build the replicas from the host state of this model
            if self._arena is not None:
                self._sync_to_host()
                self._drop_device_state()
            self._replicas = None
            tv = [v for v in self.trainable_variables if isinstance(v, Variable)]
            _, _, total = Arena.layout([v.size for v in tv], Arena.ALIGN)
            pars = Parallel.init_all(n_gpus, total)
            reps = [self] + [copy.deepcopy(self) for _ in range(n_gpus - 1)]
            for r, rep in enumerate(reps):
                rep._par = pars[r]
                rep._replicas = None
                with t.cuda.device(r):
                    rep._materialize()
            for r in range(1, n_gpus):                   # identical initial weights (they are: same host values; made explicit)
                reps[r]._arena.w.copy_(reps[0]._arena.w)
            for r in range(n_gpus):
                t.cuda.synchronize(r)
            self._replicas = reps
        # the reference draws each epoch's permutation from the GLOBAL NumPy stream (utils/MiniBatch.py:7-8): drawn
        # once here, in order, and handed to every replica (threads must not race on that stream)
        X = np.asarray(X)
        m_train = X.shape[0]
        if validation_data is None and 0. < validation_ratio < 1.:
            m_train = X.shape[0] - int(X.shape[0] * validation_ratio)
        perms = [batch_permutation(m_train, epoch, shuffle) for epoch in range(epochs)]
        errors = []

        def run(r, rep):
            try:
                t.cuda.set_device(r)
                rep.fit(X, Y, batch_size=batch_size, epochs=epochs, shuffle=shuffle, validation_data=validation_data,
                        validation_ratio=validation_ratio, verbose=verbose if r == 0 else 0, _perms=perms)
            except BaseException as e:          # noqa: B902 — re-raised in the caller's thread
                errors.append(e)
        threads = [threading.Thread(target=run, args=(r, rep), daemon=True) for r, rep in enumerate(reps)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]

    def fit(self, X, Y, batch_size=64, epochs=20, shuffle=True, validation_data=None, validation_ratio=0.1,
            draw_acc_loss=False, draw_save_path=None, verbose=1, n_gpus=None, _perms=None):
        """The reference's minibatch loop (models.py:62-118).  n_gpus=G (> 1) shards every minibatch over G GPUs of
        this box from this one process (see _fit_replicated; under torchrun each rank is a replica already and
        n_gpus is not needed).  Per replica the loop is a three-stage pipeline:
        a host thread gathers minibatch i+1.. (get_batches' permutation order) into pinned
        staging; a copy stream moves it host->device and transposes NCHW->NHWC into one of two
        device input buffers; the compute stream replays the captured step graph and copies the
        (loss, acc) pair back.  Nothing in the loop blocks on the device except ring reuse."""
        if draw_acc_loss:
            raise NotImplementedError('live matplotlib plotting is out of scope of the device path')
        if n_gpus is not None and int(n_gpus) > 1 and _perms is None:
            return self._fit_replicated(int(n_gpus), X, Y, batch_size, epochs, shuffle, validation_data, validation_ratio, verbose)
        import queue
        import threading
        t = require_cuda()
        self._materialize()
        X = np.asarray(X)
        Y = np.asarray(Y)
        if validation_data is None:
            if 0. < validation_ratio < 1.:
                split = int(X.shape[0] * validation_ratio)
                valid_X, valid_Y = X[-split:], Y[-split:]
                train_X, train_Y = X[:-split], Y[:-split]
                validation_data = (valid_X, valid_Y)
            else:
                train_X, train_Y = X, Y
        else:
            valid_X, valid_Y = validation_data
            train_X, train_Y = X, Y

        world, rank = self._par.world, self._par.rank
        m = train_X.shape[0]
        x_row = int(np.prod(train_X.shape[1:]))
        y_row = int(np.prod(train_Y.shape[1:]))
        X2 = np.ascontiguousarray(train_X, dtype=np.float32).reshape(m, x_row)     # no copy if already float32
        Y2 = np.ascontiguousarray(train_Y, dtype=np.float32).reshape(m, y_row)
        slices = batch_slices(m, batch_size)
        if not slices:
            return
        shards = []
        for a, b in slices:                         # this rank's contiguous rows of each minibatch
            rows = b - a
            shards.append((a + (rows * rank) // world, a + (rows * (rank + 1)) // world, rows))
        n_loc_max = max(hi - lo for lo, hi, _ in shards)
        if min(hi - lo for lo, hi, _ in shards) == 0:
            raise ValueError('a minibatch has fewer rows than there are ranks (%d): every rank needs at least one row of every '
                             'minibatch; use a batch size (and a sample count whose remainder) >= the number of GPUs' % world)
        res = self._fit_resources(n_loc_max, x_row, y_row, len(slices))
        depth, copy_stream, d2h_stream = res['depth'], res['copy_stream'], res['d2h_stream']
        inp = self._input_layer()
        is4d = train_X.ndim == 4 and train_X.shape[1] > 1 and train_X.shape[2] * train_X.shape[3] > 1
        x_layout = 'nhwc' if train_X.ndim == 4 else 'flat'

        # shuffle=False on float32 input: no staging copy at all — the caller's arrays are page-locked in
        # place (once, cached) and every minibatch is DMA'd straight out of them
        direct = (not shuffle) and train_X.dtype == np.float32 and train_Y.dtype == np.float32 and \
            train_X.flags.c_contiguous and train_Y.flags.c_contiguous and \
            self._pin_in_place(X2) and self._pin_in_place(Y2)

        # ---- stage 1: host gather thread --------------------------------------
        free_slots = threading.Semaphore(depth)
        ready = queue.Queue()
        perms = queue.Queue()
        stop = threading.Event()

        def gather():
            try:
                gather_loop()
            except BaseException as e:          # surface in the main thread instead of a 300 s queue timeout
                ready.put(e)

        def gather_loop():
            seq = 0
            while True:
                perm = perms.get()
                if perm is False:
                    return
                for lo, hi, _ in shards:
                    free_slots.acquire()
                    if stop.is_set():
                        return
                    slot = seq % depth
                    n = hi - lo
                    px, py = res['x'][slot].numpy(), res['y'][slot].numpy()
                    if perm is None:
                        px[:n] = X2[lo:hi]
                        py[:n] = Y2[lo:hi]
                    else:
                        np.take(X2, perm[lo:hi], axis=0, out=px[:n])
                        np.take(Y2, perm[lo:hi], axis=0, out=py[:n])
                    ready.put(slot)
                    seq += 1

        worker = threading.Thread(target=gather, daemon=True)
        if not direct:
            worker.start()
        compute_stream = t.cuda.current_stream()
        st = compute_stream.cuda_stream
        cs = copy_stream.cuda_stream
        from .layers.Embeddings import Embedding
        has_embedding = any(isinstance(l, Embedding) for l in self._forward_order())
        pending_slots = []          # (slot) whose H2D has been issued but not yet confirmed done
        step_no = 0
        try:
            for epoch in range(epochs):
                # reseeds the global RNG like the reference does (a replica thread of fit(n_gpus=...) gets the permutations drawn by its caller)
                perm = _perms[epoch] if _perms is not None else batch_permutation(m, epoch, shuffle)
                if not direct:
                    perms.put(perm)
                if verbose:
                    print('\033[0;31m Epoch[%d/%d]' % (epoch + 1, epochs))
                start_time = time.time()
                pin_metrics = res['metrics']
                for bi, (lo, hi, rows) in enumerate(shards):
                    n_loc = hi - lo
                    if direct:
                        slot, src_x, src_y = None, X2.ctypes.data + lo * x_row * 4, Y2.ctypes.data + lo * y_row * 4
                    else:
                        slot = ready.get(timeout=300)
                        if isinstance(slot, BaseException):
                            raise slot
                        src_x, src_y = res['x'][slot].data_ptr(), res['y'][slot].data_ptr()
                    j = step_no & 1
                    shape_x = (n_loc,) + train_X.shape[1:]
                    xs = inp._buf('fit_x%d' % j, shape_x, x_layout)
                    ys = inp._buf('fit_y%d' % j, (n_loc,) + train_Y.shape[1:], 'flat')
                    # ---- stage 2: copy stream ------------------------------------
                    copy_stream.wait_event(res['compute_done'][j])      # device buffer j is free again
                    if is4d:
                        stage = inp._buf('fit_x_nchw%d' % j, shape_x, 'flat')
                        lib.sn_memcpy_h2d(stage.ptr, src_x, n_loc * x_row * 4, cs)
                        lib.sn_nchw_to_nhwc(stage.ptr, xs.ptr, n_loc, shape_x[1], shape_x[2], shape_x[3], cs)
                    else:
                        lib.sn_memcpy_h2d(xs.ptr, src_x, n_loc * x_row * 4, cs)
                    lib.sn_memcpy_h2d(ys.ptr, src_y, n_loc * y_row * 4, cs)
                    res['h2d_done'][j].record(copy_stream)
                    if slot is not None:
                        res['slot_copied'][slot].record(copy_stream)
                        pending_slots.append(slot)
                    # ---- stage 3: compute stream ---------------------------------
                    compute_stream.wait_event(res['h2d_done'][j])
                    compute_stream.wait_event(res['d2h_done'][j])       # metric slot j was read back (two steps ago)
                    self._run_step(xs, ys, rows, mslot=j)
                    res['compute_done'][j].record(compute_stream)
                    # the (loss, acc) pair of this step goes home on its own stream: a DMA queued on the
                    # compute stream would sit between two graph launches
                    d2h_stream.wait_event(res['compute_done'][j])
                    lib.sn_memcpy_d2h(pin_metrics[bi].data_ptr(), self._metrics_out.data_ptr() + 8 * j, 8, d2h_stream.cuda_stream)
                    res['d2h_done'][j].record(d2h_stream)
                    step_no += 1
                    # hand staging slots whose copy has certainly finished back to the gather thread
                    while len(pending_slots) > 2:
                        res['slot_copied'][pending_slots.pop(0)].synchronize()
                        free_slots.release()
                    if validation_data is not None:
                        valid_acc, valid_loss = self.evaluate(valid_X, valid_Y, batch_size=batch_size)
                        self.valid_loss.append(valid_loss)
                        self.valid_acc.append(valid_acc)
                lib.sn_stream_sync(st)
                d2h_stream.synchronize()
                if has_embedding:
                    Embedding.check_ids()            # the reference raises IndexError on an id outside the table
                if world > 1 and self._par.own_kernels and (lib.comm_timeouts(self._par.grad.handle) + lib.comm_timeouts(self._par.stats.handle)):
                    raise RuntimeError('a data-parallel exchange timed out (a replica stopped): this epoch is invalid')
                gap = time.time() - start_time
                mets = pin_metrics.numpy()
                for bi, (_, _, rows) in enumerate(shards):
                    self.train_loss.append(float(mets[bi, 0]) / rows)
                    self.train_acc.append(float(mets[bi, 1]) / rows)
                if verbose:
                    bar = self.process_bar_trained * (self.process_bar_nums - 1) + '>'
                    msg = '\r{:d}/{:d} [{}] -{:.0f}s -{:.0f}ms/batch -batch_loss: {:.4f} -batch_acc: {:.4f} '.format(
                        m, m, bar, gap, gap * 1000 / max(1, len(slices)), self.train_loss[-1], self.train_acc[-1])
                    if validation_data is not None:
                        msg += '-val_loss: {:.4f} -val_acc: {:.4f}'.format(self.valid_loss[-1], self.valid_acc[-1])
                    print(msg)
        finally:
            stop.set()
            perms.put(False)
            for _ in range(depth + 1):
                free_slots.release()
            copy_stream.synchronize()
            d2h_stream.synchronize()
            if not direct:
                worker.join(timeout=10)

    # ---- evaluate (models.py:144-171) -------------------------------------
    def evaluate(self, X, Y, batch_size=None):
        self._materialize()
        X, Y = np.asarray(X), np.asarray(Y)
        if batch_size is not None:
            assert type(batch_size) is int
            acc_list, loss_list = [], []
            for a, b in batch_slices(X.shape[0], batch_size):
                y_hat = self.predict(X[a:b], is_training=False)
                loss, acc = self.loss.calc_loss_acc(y_hat, Y[a:b])
                acc_list.append(acc)
                loss_list.append(loss)
            # un-weighted mean of the chunk means, like the reference (:161-162)
            return sum(acc_list) / len(acc_list), sum(loss_list) / len(loss_list)
        y_hat = self.predict(X, is_training=False)
        loss, acc = self.loss.calc_loss_acc(y_hat, Y)
        return acc, loss

    # ---- save / load (models.py:208-220, 530-543) -------------------------
    def _sync_to_host(self):
        for v in self.trainable_variables:
            if isinstance(v, Variable):
                v.to_host()

    def _drop_device_state(self):
        self._arena = None
        self._par = None
        self._graphs = {}
        self._warm = set()
        self._metrics_out = None
        self._fit_res = None
        self._cap_stream = None
        self._comm_stream = None
        self._bf16_shadow = None
        self._prep_tasks = {}
        self._split = None
        self._replicas = None
        if isinstance(self.trainable_variables, VariableList):      # drops the handle on the old arena
            self.trainable_variables = list(self.trainable_variables)
        for ptr in list(getattr(self, '_pinned_user', {})):
            try:
                lib.sn_host_unregister(ptr)
            except Exception:
                pass
        self._pinned_user = {}
        for layer in self._forward_order():
            layer._bufs = {}
            for v in layer.variables:
                if isinstance(v, Variable):
                    v._dev = None
                    v._grads = None


class Sequential(_Trainer):
    def __init__(self, layers=None):
        self.layers = [] if layers is None else layers
        self._init_histories()

    def add(self, layer):
        self.layers.append(layer)

    def compile(self, optimizer, loss, precision='fp32', fuse=True, sync_bn=None):
        """models.py:30-43: connect the layers in order (shape inference + host-side parameter
        creation, same RNG consumption as the reference), collect the trainable variables."""
        assert self.layers
        trainable_variables = []
        next_layer = None
        for layer in self.layers:
            layer.connect(next_layer)
            next_layer = layer
            for var in layer.variables:
                if var.require_grads:
                    trainable_variables.append(var)
        self.trainable_variables = trainable_variables
        self._finish_compile(optimizer, loss, precision, fuse, sync_bn)

    def _forward_order(self):
        return self.layers

    def _backward_order(self):
        return list(reversed(self.layers))

    def _input_layer(self):
        return self.layers[0]

    def _output_layer(self):
        return self.layers[-1]

    def pop(self, index=-1):
        layer = self.layers.pop(index)
        print('success delete %s layer' % (layer.__class__.__name__))

    def save(self, save_path):
        self._sync_to_host()
        with open(save_path + '.pkl', 'wb') as f:
            pickle.dump([self.layers, self.optimizer, self.loss], f)

    def load(self, model_path):
        with open(model_path + '.pkl', 'rb') as f:
            layers, optimizer, loss = pickle.load(f)
        self.layers, self.optimizer, self.loss = layers, optimizer, loss
        self.trainable_variables = [v for l in layers for v in l.variables if v.require_grads]
        self._drop_device_state()
        if layers:
            self.precision = layers[0].precision


class Model(_Trainer):
    def __init__(self, inputs=None, outputs=None):
        self.inputs = inputs
        self.outputs = outputs
        self._init_histories()

    @staticmethod
    def _kahn(start, edges_out, edges_in):
        """Kahn's algorithm from `start` following `edges_out` (models.py:287-356)."""
        seen, order, indeg = {}, [], {}
        queue = [start]
        while queue:
            n = queue.pop(0)
            if id(n) in seen:
                continue
            seen[id(n)] = n
            for m in edges_out(n):
                queue.append(m)
        for n in seen.values():
            indeg[id(n)] = sum(1 for p in edges_in(n) if id(p) in seen)
        ready = [n for n in seen.values() if indeg[id(n)] == 0]
        while ready:
            n = ready.pop(0)
            order.append(n)
            for m in edges_out(n):
                indeg[id(m)] -= 1
                if indeg[id(m)] == 0:
                    ready.append(m)
        return order

    def compile(self, optimizer, loss, precision='fp32', fuse=True, sync_bn=None):
        assert self.inputs is not None and self.outputs is not None
        self.forward_graph = self._kahn(self.inputs, lambda n: n.outbound_layers, lambda n: n.inbounds)
        self.backward_graph = self._kahn(self.outputs, lambda n: n.inbounds, lambda n: n.outbound_layers)
        tv = []
        for layer in self.forward_graph:
            for var in layer.variables:
                # only real Variables: the reference also registers the Layer operands of Add
                # here (models.py:304-307), which then breaks its optimizer
                if isinstance(var, Variable) and var.require_grads and var not in tv:
                    tv.append(var)
        self.trainable_variables = tv
        self._finish_compile(optimizer, loss, precision, fuse, sync_bn)

    def _forward_order(self):
        return self.forward_graph

    def _backward_order(self):
        return self.backward_graph

    def _input_layer(self):
        return self.inputs

    def _output_layer(self):
        return self.outputs

    def save(self, save_path):
        self._sync_to_host()
        with open(save_path + '.pkl', 'wb') as f:
            pickle.dump([self.forward_graph, self.backward_graph, self.optimizer, self.loss], f)

    def load(self, model_path):
        with open(model_path + '.pkl', 'rb') as f:
            fg, bg, optimizer, loss = pickle.load(f)
        self.forward_graph, self.backward_graph, self.optimizer, self.loss = fg, bg, optimizer, loss
        self.inputs, self.outputs = fg[0], bg[0]
        self.trainable_variables = [v for l in fg for v in l.variables if isinstance(v, Variable) and v.require_grads]
        self._drop_device_state()
