from .transform_convert import axisangle2mat, mat2axisangle, Axisangle2MatFunction, Mat2AxisangleFunction
from .transform import (
    RigidTransform,
    mat_first2last,
    mat_last2first,
    ax_first2last,
    ax_last2first,
    mat_update_resolution,
    ax_update_resolution,
    mat_transform_points,
    ax_transform_points,
    transform_points,
)
