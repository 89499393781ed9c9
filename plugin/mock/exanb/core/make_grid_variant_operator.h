// mock of exanb/core/make_grid_variant_operator.h: instantiates the operator template for one field set
#pragma once
#include <memory>
#include <exanb/core/grid.h>
#include <onika/scg/operator.h>
namespace exanb {
template<template<class> class OperatorTmpl>
inline std::shared_ptr<onika::scg::OperatorNode> make_grid_variant_operator() { return std::make_shared< OperatorTmpl< GridMock<> > >(); }
template<class OperatorT>
inline std::shared_ptr<onika::scg::OperatorNode> make_simple_operator() { return std::make_shared<OperatorT>(); }
}
