"""Minimal hand-written encoder of TensorFlow GraphDef protobufs (test infrastructure).

TensorFlow itself cannot be installed here, but OpenCV's `cv2.dnn.readNetFromTensorflow` - an independent engine written to
reproduce frozen TensorFlow graphs - is in the image.  This module builds the few GraphDef messages the pin tests need
(Placeholder, Const, Conv2D, BiasAdd, Relu, MaxPool, ResizeBilinear, FusedBatchNorm) directly in the protobuf wire format.
Field numbers are those of tensorflow/core/framework/{graph,node_def,attr_value,tensor,tensor_shape,types}.proto:
  GraphDef.node = 1;  NodeDef: name 1, op 2, input 3, attr 5 (map<string, AttrValue>: key 1, value 2)
  AttrValue: list 1, s 2, i 3, f 4, b 5, type 6, shape 7, tensor 8;  ListValue.i = 3 (packed)
  TensorProto: dtype 1, tensor_shape 2, tensor_content 4;  TensorShapeProto.dim = 2 {size 1};  DT_FLOAT = 1, DT_INT32 = 3
"""
import struct

import numpy as np

DT = {np.dtype('float32'): 1, np.dtype('int32'): 3}


def _varint(n):
    n &= (1 << 64) - 1
    out = b''
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out += bytes([b | 0x80])
        else:
            return out + bytes([b])


def _ld(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _vi(field, n):
    return _varint(field << 3) + _varint(n)


def _shape(dims):
    return b''.join(_ld(2, _vi(1, d)) for d in dims)


def _tensor(arr):
    arr = np.ascontiguousarray(arr)
    return _vi(1, DT[arr.dtype]) + _ld(2, _shape(arr.shape)) + _ld(4, arr.tobytes())


def attr(name, value):
    return _ld(5, _ld(1, name.encode()) + _ld(2, value))


def a_type(t=1):
    return _vi(6, t)


def a_str(s):
    return _ld(2, s.encode())


def a_bool(b):
    return _vi(5, int(b))


def a_float(f):
    return _varint((4 << 3) | 5) + struct.pack('<f', f)


def a_ints(xs):
    return _ld(1, _ld(3, b''.join(_varint(x) for x in xs)))


def a_shape(dims):
    return _ld(7, _shape(dims))


def node(name, op, inputs=(), attrs=()):
    return _ld(1, _ld(1, name.encode()) + _ld(2, op.encode()) + b''.join(_ld(3, i.encode()) for i in inputs) + b''.join(attrs))


def placeholder(name, shape):
    return node(name, 'Placeholder', (), [attr('dtype', a_type(1)), attr('shape', a_shape(list(shape)))])


def const(name, arr):
    arr = np.asarray(arr)
    return node(name, 'Const', (), [attr('dtype', a_type(DT[arr.dtype])), attr('value', _ld(8, _tensor(arr)))])


def conv2d(name, x, w, stride, padding='SAME'):
    return node(name, 'Conv2D', (x, w), [attr('T', a_type()), attr('strides', a_ints([1, stride, stride, 1])), attr('padding', a_str(padding)),
                                         attr('data_format', a_str('NHWC')), attr('dilations', a_ints([1, 1, 1, 1]))])


def bias_add(name, x, b):
    return node(name, 'BiasAdd', (x, b), [attr('T', a_type()), attr('data_format', a_str('NHWC'))])


def relu(name, x):
    return node(name, 'Relu', (x,), [attr('T', a_type())])


def max_pool(name, x, size=2, stride=2, padding='SAME'):
    return node(name, 'MaxPool', (x,), [attr('T', a_type()), attr('ksize', a_ints([1, size, size, 1])), attr('strides', a_ints([1, stride, stride, 1])),
                                        attr('padding', a_str(padding)), attr('data_format', a_str('NHWC'))])


def resize_bilinear(name, x, size_const, align_corners=False):
    return node(name, 'ResizeBilinear', (x, size_const), [attr('T', a_type()), attr('align_corners', a_bool(align_corners)),
                                                          attr('half_pixel_centers', a_bool(False))])


def fused_batch_norm(name, x, gamma, beta, mean, var, epsilon):
    return node(name, 'FusedBatchNorm', (x, gamma, beta, mean, var), [attr('T', a_type()), attr('epsilon', a_float(epsilon)),
                                                                       attr('is_training', a_bool(False)), attr('data_format', a_str('NHWC'))])


def run_opencv(graph_bytes, x_nhwc):
    """Runs the frozen graph through OpenCV's TensorFlow importer; NHWC in, NHWC out (float32)."""
    import cv2
    net = cv2.dnn.readNetFromTensorflow(np.frombuffer(graph_bytes, np.uint8))
    net.setInput(np.ascontiguousarray(np.asarray(x_nhwc, dtype=np.float32).transpose(0, 3, 1, 2)))
    return net.forward().transpose(0, 2, 3, 1)
