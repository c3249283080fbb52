"""BatchNormalization on the device (reference: layers/Normalization.py:8-230)."""
import numpy as np

from .Base import Layer, Variable, flush_deferred
from ..device import DeviceTensor, current_stream, zeros, torch
from .._lib import lib, POOL_MODE, ACT
from ..utils.Initializers import get_initializer


class BatchNormalization(Layer):
    def __init__(self, epsilon=1e-6, momentum=0.99, axis=1, gamma_initializer='ones', beta_initializer='zeros',
                 moving_mean_initializer='zeros', moving_variance_initializer='ones'):
        self.epsilon = epsilon
        self.axis = axis
        self.momentum = momentum
        self.gamma_initializer = get_initializer(gamma_initializer)
        self.beta_initializer = get_initializer(beta_initializer)
        self.moving_mean_initializer = get_initializer(moving_mean_initializer)
        self.moving_variance_initializer = get_initializer(moving_variance_initializer)
        super(BatchNormalization, self).__init__()

    def _infer(self, shape):
        assert len(shape) >= 2
        if len(shape) not in (2, 4) or self.axis not in (1, -1 if len(shape) == 2 else 1):
            raise NotImplementedError('device BatchNormalization normalises axis 1 of (N,C,H,W) or (N,M) inputs')
        n_in = shape[1]
        self.variables.append(Variable(self.gamma_initializer(n_in), name='bn_gamma'))
        self.variables.append(Variable(self.beta_initializer(n_in), name='bn_beta'))
        self._mm_host = np.asarray(self.moving_mean_initializer(n_in), dtype=np.float64)
        self._mv_host = np.asarray(self.moving_variance_initializer(n_in), dtype=np.float64)
        self._stats = None
        self._fuse_with_pool = None      # set by the model when a 2x2/stride-2 MaxPooling2D is the only consumer
        self._fused_active = False
        self.output_shape = shape

    def connect(self, prev_layer):
        self._infer(prev_layer.output_shape)
        Layer.connect(self, prev_layer)

    def __call__(self, prev_layer):
        super(BatchNormalization, self).__call__(prev_layer)
        self._infer(self.input_shape)
        return self

    def _takes_bf16_map(self, shape):
        """Will forward() run fused with the 2x2 pool on an input of this (N, C, H, W) shape?  Then the
        producing convolution may hand its output map over in bf16 (bn_pool kernels read either type)."""
        pool = self._fuse_with_pool
        if pool is None or len(shape) != 4:
            return False
        N, C, H, W = shape
        return bool(lib.load().sn_bn_pool_supported(N, H, W, C)) and C % 4 == 0 and (C // 4) & (C // 4 - 1) == 0 and C <= 1024

    def _device_stats(self, C):
        if self._stats is None:
            self._stats = dict(mm=DeviceTensor.from_numpy(self._mm_host), mv=DeviceTensor.from_numpy(self._mv_host),
                               mean=DeviceTensor.empty((C,)), invstd=DeviceTensor.empty((C,)),
                               scratch=zeros(28 * C + 2, torch().float64))
        return self._stats

    # ---- cross-replica statistics (compile(sync_bn=True) under torchrun) ---------------------
    def _sync(self, is_training=True):
        """The data-parallel context (comm.Parallel) when this layer must sum its statistics over all
        replicas, else None.  `_par` is set by the model when it places the parameters, `_global_batch`
        before every step."""
        if not (is_training and getattr(self, '_sync_bn', False)):
            return None
        par = getattr(self, '_par', None)
        if par is None or par.world == 1:
            return None
        return par

    def _global_rows(self, x, rows):
        gb = getattr(self, '_global_batch', None)
        n_loc = x.shape[0]
        return rows if not gb else (rows // n_loc) * int(gb)

    def _finalize_stats(self, par, s, gamma, beta, rows_g, C, st):
        """Batch statistics -> apply constants.  Data parallel: the cross-replica sum rides inside the
        finalize kernel (peer-mapped memory, csrc/bn.cu bn_finalize_dp_kernel); only the NCCL fallback
        all-reduces the scratch first."""
        if par is not None and par.stats is not None:
            lib.sn_batchnorm_finalize_stats_dp(par.stats.handle, gamma.output_tensor.ptr, beta.output_tensor.ptr, s['mean'].ptr,
                                               s['invstd'].ptr, s['mm'].ptr, s['mv'].ptr, s['scratch'].data_ptr(), rows_g, C,
                                               float(self.epsilon), float(self.momentum), st)
            return
        if par is not None:
            par.allreduce_stats(s['scratch'][:16 * C], st)
        lib.sn_batchnorm_finalize_stats(gamma.output_tensor.ptr, beta.output_tensor.ptr, s['mean'].ptr, s['invstd'].ptr,
                                        s['mm'].ptr, s['mv'].ptr, s['scratch'].data_ptr(), rows_g, C,
                                        float(self.epsilon), float(self.momentum), st)

    def _finalize_bwd(self, par, s, gamma, rows_g, C, st):
        if par is not None and par.stats is not None and __import__('os').environ.get('SN_BN_SPLIT_EXCHANGE', '1') == '1':
            # split exchange: send the local sums, run whatever the layers behind us postponed (a Conv2D's weight
            # gradient: ~30 us of independent work that absorbs the NVLink latency and the skew between the
            # replicas), then poll + finalise
            lib.sn_batchnorm_dp_push(par.stats.handle, s['scratch'].data_ptr(), C, st)
            flush_deferred()
            lib.sn_batchnorm_finalize_bwd_dp_pushed(par.stats.handle, gamma.output_tensor.ptr, s['mean'].ptr, s['invstd'].ptr,
                                                    s['scratch'].data_ptr(), rows_g, C, st)
            return
        if par is not None and par.stats is not None:
            lib.sn_batchnorm_finalize_bwd_dp(par.stats.handle, gamma.output_tensor.ptr, s['mean'].ptr, s['invstd'].ptr,
                                             s['scratch'].data_ptr(), rows_g, C, st)
            return
        if par is not None:
            par.allreduce_stats(s['scratch'][:16 * C], st)
        lib.sn_batchnorm_finalize_bwd(gamma.output_tensor.ptr, s['mean'].ptr, s['invstd'].ptr, s['scratch'].data_ptr(), rows_g, C, st)

    @property
    def moving_mean(self):
        return self._stats['mm'].numpy() if self._stats is not None else self._mm_host

    @property
    def moving_variance(self):
        return self._stats['mv'].numpy() if self._stats is not None else self._mv_host

    def forward(self, is_training=True):
        """layers/Normalization.py:93-156 (biased variance, eps inside the sqrt)."""
        gamma, beta = self._param_ptrs()
        x = self.input_tensor
        self.input_shape = x.shape
        C = x.shape[1]
        rows = x.size // C
        s = self._device_stats(C)
        st = current_stream()
        pool = self._fuse_with_pool
        # the producing convolution already left sum(x), sum(x^2) in our scratch (its epilogue): no reduce pass
        prestats = bool(is_training and getattr(self, '_stats_from', None) is not None
                        and self._stats_from is self.inbounds[0])
        self._stats_from = None
        sync = self._sync(is_training)
        rows_g = self._global_rows(x, rows) if sync is not None else rows
        if sync is not None and not (C % 4 == 0 and (C // 4) & (C // 4 - 1) == 0 and C <= 1024):
            raise NotImplementedError('sync_bn needs a channel count of the form 4*2^k <= 1024 (got %d)' % C)
        self._fused_active = bool(is_training and pool is not None and x.ndim == 4 and
                                  lib.load().sn_bn_pool_supported(x.shape[0], x.shape[2], x.shape[3], C))
        if self._fused_active:
            # BN + 2x2 max-pool in one pass (csrc/bn_pool.cu): the normalised map is never materialised,
            # so this layer has no output_tensor of its own in this mode (compile(..., fuse=False) keeps it).
            N, _, H, W = x.shape
            yp, code = pool._fused_buffers(N, C, H, W)
            # the next convolution's bf16 operand copy, written while the pooled map is produced
            nxt = pool.outbound_layers[0] if len(pool.outbound_layers) == 1 else None
            ypb = nxt._bf16_input_buffer(yp.shape) if hasattr(nxt, '_bf16_input_buffer') else None
            ypb_ptr = ypb.data_ptr() if ypb is not None else None
            if ypb is not None:
                nxt._xb_filled_for = yp
            if prestats or sync is not None:
                if not prestats:
                    lib.sn_batchnorm_local_stats(x.ptr, s['scratch'].data_ptr(), rows, C, st)
                self._finalize_stats(sync, s, gamma, beta, rows_g, C, st)
                fwd_apply = lib.sn_bn_pool_fwd_apply_xbf16 if x.dtype == 'bf16' else lib.sn_bn_pool_fwd_apply
                fwd_apply(x.ptr, yp.ptr, ypb_ptr, code.data_ptr(), s['scratch'].data_ptr(), N, H, W, C, POOL_MODE[pool.mode], st)
            else:
                assert x.dtype == 'f32'         # a bf16 map always comes with the producer's statistics
                lib.sn_bn_pool_fwd_train(x.ptr, gamma.output_tensor.ptr, beta.output_tensor.ptr, yp.ptr, ypb_ptr, code.data_ptr(),
                                         s['mean'].ptr, s['invstd'].ptr, s['mm'].ptr, s['mv'].ptr, s['scratch'].data_ptr(),
                                         N, H, W, C, float(self.epsilon), float(self.momentum), POOL_MODE[pool.mode], st)
            pool._fused_output = yp
            pool.input_tensor = None
            self.output_tensor = None
            self._grads_written = False
            self._grads_premasked = False
            return
        y = self._buf('out', x.shape, x.layout)
        relu_out = getattr(self, '_relu_out', None) is not None
        if is_training and (prestats or sync is not None):
            if not prestats:
                lib.sn_batchnorm_local_stats(x.ptr, s['scratch'].data_ptr(), rows, C, st)
            self._finalize_stats(sync, s, gamma, beta, rows_g, C, st)
            # the Activation('relu') behind this layer, when compile() folded it in, rides on the apply pass
            lib.sn_batchnorm_fwd_apply(x.ptr, y.ptr, s['scratch'].data_ptr(), rows, C, 1 if relu_out else 0, st)
            relu_out = False
        elif is_training:
            lib.sn_batchnorm_fwd_train(x.ptr, gamma.output_tensor.ptr, beta.output_tensor.ptr, y.ptr, s['mean'].ptr,
                                       s['invstd'].ptr, s['mm'].ptr, s['mv'].ptr, s['scratch'].data_ptr(), rows, C,
                                       float(self.epsilon), float(self.momentum), st)
        else:
            lib.sn_batchnorm_fwd_infer(x.ptr, gamma.output_tensor.ptr, beta.output_tensor.ptr, y.ptr, s['mm'].ptr,
                                       s['mv'].ptr, rows, C, float(self.epsilon), st)
        if relu_out:
            lib.sn_relu_fwd(y.ptr, y.ptr, y.size, st)          # paths without the flag: in place
        self.output_tensor = y
        super(BatchNormalization, self).forward(is_training)

    def backward(self):
        """layers/Normalization.py:186-230.  If the single producer is a ReLU-fused layer whose
        only consumer is this BN, its mask `pre-activation < 0` (utils/Activator.py:25) is applied
        while dX is written, saving that layer one read-modify-write pass over its gradient."""
        gamma, beta = self.variables
        x = self.input_tensor
        C = x.shape[1]
        s = self._stats
        st = current_stream()
        sync = self._sync(True)
        rows_g = self._global_rows(x, x.size // C) if sync is not None else x.size // C
        if self._fused_active:
            pool = self._fuse_with_pool
            N, _, H, W = x.shape
            for layer in self.inbounds:
                fuse = bool(layer.require_grads and layer._fused_relu() and len(layer.outbound_layers) == 1
                            and not layer._grads_written)

                # what a Conv2D producer whose gradient map is exactly our dX can take directly:
                # its bias gradient (column sums of dX, post-mask), and in the bf16 tier the bf16 copy
                # of dX its wgrad / dgrad read (then the fp32 map need not be written at all)
                sole = bool(layer.require_grads and len(layer.outbound_layers) == 1 and not layer._grads_written
                            and hasattr(layer, '_bf16_grad_plan') and (fuse or layer._act_code() == ACT['linear']))
                colsum = gb = None
                want_f32 = True
                if sole:
                    bias = layer.variables[1]
                    if bias.require_grads:
                        colsum = bias.grads.ptr
                        layer._db_done = True
                    wg16, dg16 = layer._bf16_grad_plan()
                    if wg16 or dg16:
                        gb = layer._bf16_buf('dy_bf16', x.size).data_ptr()
                        layer._gb_ready = True
                        needs_dx = any(l.require_grads for l in layer.inbounds)
                        want_f32 = not (wg16 and (dg16 or not needs_dx))

                def fused_writer(dst, acc, fuse=fuse, colsum=colsum, gb=gb, want_f32=want_f32):
                    out = self._buf('dx_tmp', x.shape, x.layout) if acc else dst      # (fp32 whatever the type of x)
                    dgam = gamma.grads.ptr if gamma.require_grads else None
                    dbet = beta.grads.ptr if beta.require_grads else None
                    xb = x.dtype == 'bf16'
                    if sync is not None:
                        # local sums (+ this replica's dgamma / dbeta), all-reduce, global constants, apply
                        (lib.sn_bn_pool_bwd_local_xbf16 if xb else lib.sn_bn_pool_bwd_local)(pool.grads.ptr, pool._code.data_ptr(), x.ptr, pool.output_tensor.ptr,
                                                 gamma.output_tensor.ptr, beta.output_tensor.ptr, s['mean'].ptr, s['invstd'].ptr,
                                                 dgam, dbet, s['scratch'].data_ptr(), N, H, W, C, POOL_MODE[pool._code_mode], st)
                        self._finalize_bwd(sync, s, gamma, rows_g, C, st)
                        (lib.sn_bn_pool_bwd_apply_xbf16 if xb else lib.sn_bn_pool_bwd_apply)(pool.grads.ptr, pool._code.data_ptr(), x.ptr,
                                                 out.ptr if (want_f32 or acc) else None, gb, colsum, s['scratch'].data_ptr(),
                                                 N, H, W, C, POOL_MODE[pool._code_mode], 1 if fuse else 0, st)
                    else:
                        (lib.sn_bn_pool_bwd_xbf16 if xb else lib.sn_bn_pool_bwd)(pool.grads.ptr, pool._code.data_ptr(), x.ptr, pool.output_tensor.ptr,
                                           gamma.output_tensor.ptr, beta.output_tensor.ptr, s['mean'].ptr,
                                           s['invstd'].ptr, out.ptr if (want_f32 or acc) else None, gb, colsum, dgam, dbet,
                                           s['scratch'].data_ptr(), N, H, W, C, POOL_MODE[pool._code_mode], 1 if fuse else 0, st)
                    if acc:
                        dst.add_(out)
                if layer.require_grads:
                    self._push_grad(layer, fused_writer)
                    layer._grads_premasked = fuse
                else:
                    fused_writer(self._buf('dx_unused', x.shape, x.layout), 0, False)
            return
        g = self.grads
        rows = x.size // C
        for layer in self.inbounds:
            fuse = bool(layer.require_grads and getattr(layer, '_fused_relu', lambda: False)()
                        and len(layer.outbound_layers) == 1 and not layer._grads_written)

            def writer(dst, acc, fuse=fuse):
                dgam = gamma.grads.ptr if gamma.require_grads else None
                dbet = beta.grads.ptr if beta.require_grads else None
                if sync is not None:
                    lib.sn_batchnorm_bwd_local(g.ptr, x.ptr, s['mean'].ptr, s['invstd'].ptr, dgam, dbet,
                                               s['scratch'].data_ptr(), rows, C, st)
                    self._finalize_bwd(sync, s, gamma, rows_g, C, st)
                    lib.sn_batchnorm_bwd_apply(g.ptr, x.ptr, dst.ptr, s['scratch'].data_ptr(), rows, C, acc, 1 if fuse else 0, st)
                    return
                lib.sn_batchnorm_bwd(g.ptr, x.ptr, gamma.output_tensor.ptr, s['mean'].ptr, s['invstd'].ptr, dst.ptr,
                                     dgam, dbet, s['scratch'].data_ptr(), rows, C, acc, 1 if fuse else 0, st)
            if layer.require_grads:
                self._push_grad(layer, writer)
                layer._grads_premasked = fuse
            else:
                # parameter gradients are still needed when nothing upstream wants dX
                tmp = self._buf('dx_unused', x.shape, x.layout)
                writer(tmp, 0, False)

    def __getstate__(self):
        if self._stats is not None:
            self._mm_host, self._mv_host = self._stats['mm'].numpy(), self._stats['mv'].numpy()
        d = Layer.__getstate__(self)
        d['_stats'] = None
        return d
