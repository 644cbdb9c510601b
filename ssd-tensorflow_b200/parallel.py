"""Data parallelism over the GPUs of one box: one process per GPU, images sharded by rank,
ONE exchange step -- a sum all-reduce (NCCL over NVLink / NVSwitch through torch.distributed)
on the flat gradient buffer -- then the fused Momentum update with the 1/world scale folded in.

The reference has no multi-GPU path.  The multibox loss normalises per
image and takes a batch mean (ssdvgg.py:513-521,552-560), so the rank-local gradient of the local
mean, summed over ranks and divided by the world size, equals the single-GPU gradient of the
concatenated batch up to summation order; the L2 term is replica-identical and is applied once,
after the reduce, inside the update kernel.
"""
import numpy as np


def shard_range(global_count, rank, world):
    """Contiguous shard [lo, hi) of `global_count` items for `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(global_count), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DeviceArray:
    """__cuda_array_interface__ view of a raw device pointer (zero copy into torch)."""
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {'shape': (int(count),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 3}


def allreduce_buckets(tensor, buckets, world, group=None, before=None):
    """Sum all-reduce of `tensor` bucket by bucket, in the given order (the order in which the backward finishes them).
    `before(k)` is called right before bucket k is issued (the CUDA path makes the side stream wait for that bucket's
    gradient-ready event there).  Returns the scale the caller folds into the update (1 / world)."""
    import torch.distributed as dist
    if world > 1:
        for k, (lo, hi) in enumerate(buckets):
            if before is not None:
                before(k)
            dist.all_reduce(tensor[lo:hi], op=dist.ReduceOp.SUM, group=group)
    return 1.0 / float(world)


def average_gradients(tensor, world, group=None):
    """all-reduce(sum) in place; the caller folds 1/world into the update (returns that scale)."""
    import torch.distributed as dist
    if world > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / float(world)


class DataParallelTrainer:
    """Wraps one ssdb.Net per process.  step() = forward + loss + backward on the local shard,
    gradient all-reduce, fused update.  Equal local batch sizes are assumed (weak scaling)."""

    def __init__(self, net, group=None):
        import torch
        import torch.distributed as dist
        import ssdb
        self.net = net
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        ptr, count = net.flat_buffer(ssdb.GRAD)
        self.grads = torch.as_tensor(_DeviceArray(ptr, count), device='cuda')
        pptr, _ = net.flat_buffer(ssdb.PARAM)
        self.params = torch.as_tensor(_DeviceArray(pptr, count), device='cuda')
        self.buckets = net.grad_buckets()
        self.side = torch.cuda.Stream() if self.world > 1 else None

    def broadcast_parameters(self, src=0):
        import torch.distributed as dist
        if self.world > 1:
            dist.broadcast(self.params, src=src, group=self.group)

    def _step_host(self, lr, momentum, weight_decay, **feed):
        import torch
        if self.world == 1:
            if 'labels' in feed:
                return self.net.train_step_host(feed['images'], feed['labels'], lr, momentum, weight_decay)
            res, losses, _ = self.net.train_step_host_gt(feed['images'], feed['gt'], feed['gt_count'], lr, momentum, weight_decay)
            return res, losses
        # everything of the local step is enqueued on the engine's streams without waiting; the all-reduce of each gradient
        # bucket and then the fused update run on the side stream as the buckets complete, beside the remaining backward
        res = self.net.train_step_host_begin(weight_decay=weight_decay, **feed)
        with torch.cuda.stream(self.side):
            scale = allreduce_buckets(self.grads, self.buckets, self.world, self.group,
                                      before=lambda k: self.net.wait_grad_bucket(k, self.side.cuda_stream))
            self.net.apply_update(lr, momentum, weight_decay, grad_post_scale=scale, stream=self.side.cuda_stream)
        losses = self.net.train_step_host_end()
        self.side.synchronize()
        return res, losses

    def step_host(self, images, labels, lr, momentum, weight_decay):
        """The data-parallel equivalent of sess.run([net.result, net.losses, net.optimizer], feed_dict): host arrays in,
        (result, losses) out; copies overlap the compute inside the library, the all-reduce overlaps the backward."""
        return self._step_host(lr, momentum, weight_decay, images=images, labels=labels)

    def step_host_gt(self, images, gt, gt_count, lr, momentum, weight_decay):
        """step_host with raw ground truth [B,G,5] + counts instead of dense labels (fused anchor matching)."""
        return self._step_host(lr, momentum, weight_decay, images=images, gt=gt, gt_count=gt_count)

    def step(self, images_ptr, labels_ptr, local_batch, lr, momentum, weight_decay, losses_ptr=None, result_ptr=None,
             gt_ptr=None, gt_count_ptr=None, G=0):
        import torch
        st = torch.cuda.current_stream().cuda_stream
        self.net.train_step(images_ptr, local_batch, labels_ptr=labels_ptr, gt_ptr=gt_ptr, gt_count_ptr=gt_count_ptr, G=G,
                            lr=lr, momentum=momentum, weight_decay=weight_decay, grad_scale=1.0, apply_update=False,
                            losses_ptr=losses_ptr, result_ptr=result_ptr, stream=st)
        if self.world > 1:
            # the whole step is enqueued; the all-reduce of each gradient bucket starts on a side stream as soon as the
            # backward has finished that range (heads / deep layers first) and overlaps the remaining dgrad / wgrad kernels;
            # only the last, small bucket (conv1_1 .. conv3_3, 7 MB) is reduced after the backward
            with torch.cuda.stream(self.side):
                scale = allreduce_buckets(self.grads, self.buckets, self.world, self.group,
                                          before=lambda k: self.net.wait_grad_bucket(k, self.side.cuda_stream))
            torch.cuda.current_stream().wait_stream(self.side)
        else:
            scale = 1.0
        self.net.apply_update(lr, momentum, weight_decay, grad_post_scale=scale, stream=st)
