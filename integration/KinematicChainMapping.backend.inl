// KinematicChainMapping.backend.inl — the code integration/apply_backend_patch.py splices into the reference's
// src/sofa/RigidBodyDynamics/KinematicChainMapping.{h,inl} so that the component calls the B200 back end
// (include/rbd_b200.h) instead of pinocchio.  SOURCE ONLY: it uses SOFA and pinocchio types and cannot be compiled in
// this repository's image; everything SOFA-free it relies on is in rbd_backend_glue.h, which IS compiled and tested
// here (tests/test_integration_patch.py).  Sections are delimited by "//@@ name" lines; the applier replaces function
// bodies found by name and parameter list (brace matching), it never needs a line of the reference's own text.
//
// Batch dimension: the component serves N robots at once when its state vectors hold N robots back to back — input1
// N * nq_joints Vec1 dofs, input2 N root poses, output N * (njoints + extra frames) rigid frames — which is exactly
// the instance-major layout of the *_host entry points.  N = 1 is today's scene; nothing else changes.

//@@ h.include
#include <sofa/RigidBodyDynamics/rbd_backend_glue.h>

//@@ h.members
    // ---- B200 back end (replaces m_data: the pinocchio::Data workspace) ----
    rbd_glue::Backend m_rbd;                       ///< model + workspace handles of the C ABI
    rbd_glue::RowsCSR m_rbdRows;                   ///< scratch of the MatrixDeriv overload (grow-only)
    Data<int> d_device;                            ///< CUDA device the robots of this component live on
    Data<unsigned> d_robots;                       ///< robots stored back to back in the state vectors (default 1)
    void stageBackend();                           ///< flattens m_model + frame tables, creates the handles
    void fail(const char* what);                   ///< msg_error + ComponentState::Invalid (the reference's convention)
    sofa::Size robots(sofa::Size in1Size) const;   ///< batch size implied by the size of input1

//@@ inl.ctor_init
      , d_device(initData(&d_device, 0, "device", "CUDA device index of the rbd_b200 workspace"))
      , d_robots(initData(&d_robots, 1u, "robots", "Number of robots stored back to back in the state vectors"))

//@@ inl.helpers
  namespace
  {
    // axis of pinocchio's unaligned joint models (the aligned ones carry it in their type)
    inline void axisOf(const pinocchio::JointModel &jm, double *axis)
    {
      if (const auto *ru = boost::get<pinocchio::JointModelRevoluteUnaligned>(&jm.toVariant()))
        for (int k = 0; k < 3; ++k) axis[k] = ru->axis[k];
      else if (const auto *pu = boost::get<pinocchio::JointModelPrismaticUnaligned>(&jm.toVariant()))
        for (int k = 0; k < 3; ++k) axis[k] = pu->axis[k];
      else if (const auto *rub = boost::get<pinocchio::JointModelRevoluteUnboundedUnaligned>(&jm.toVariant()))
        for (int k = 0; k < 3; ++k) axis[k] = rub->axis[k];
    }
  }

  template <class TIn, class TInRoot, class TOut>
  void KinematicChainMapping<TIn, TInRoot, TOut>::fail(const char *what)
  {
    msg_error() << "rbd_b200 " << what << ": " << rbd_last_error();
    d_componentState.setValue(sofa::core::objectmodel::ComponentState::Invalid);
  }

  template <class TIn, class TInRoot, class TOut>
  sofa::Size KinematicChainMapping<TIn, TInRoot, TOut>::robots(sofa::Size in1Size) const
  {
    const sofa::Size perRobot = static_cast<sofa::Size>(m_rbd.info.nq_joints > 0 ? m_rbd.info.nq_joints : 1);
    return in1Size >= perRobot ? in1Size / perRobot : 1;
  }

  // Called at the end of init(): by then the loader has handed over the model and both frame tables
  // (URDFModelLoader.cpp:263-268).  Everything derived from the model is built once here, nothing per step.
  template <class TIn, class TInRoot, class TOut>
  void KinematicChainMapping<TIn, TInRoot, TOut>::stageBackend()
  {
    if (!m_model)
      return;
    rbd_glue::Descriptor desc;
    std::string why;
    if (!rbd_glue::fillDescriptor(*m_model, m_bodyCoMFrames, m_extraFrames, axisOf, &desc, &why))
    {
      msg_error() << why;
      d_componentState.setValue(sofa::core::objectmodel::ComponentState::Invalid);
      return;
    }
    const unsigned n = d_robots.getValue() > 0u ? d_robots.getValue() : 1u;
    if (m_rbd.create(desc.d, d_device.getValue(), static_cast<int64_t>(n)) != RBD_OK)
      fail("model / workspace creation");
  }

//@@ inl.init_tail
    stageBackend();

//@@ inl.setModel.body
    assert(model);
    m_model = model;
    // no pinocchio::Data: the per-instance workspace lives in the back end (stageBackend, called from init())

//@@ inl.apply.body
    assert(m_model);

    if (d_componentState.getValue() == sofa::core::objectmodel::ComponentState::Invalid)
      return;

    // VecCoord<Vec1> / VecCoord<Rigid3> are contiguous arrays of 1 / 7 doubles (Conversions.h:58-71 documents the
    // field order): N robots back to back are the instance-major arrays rbd_apply_host takes
    const sofa::Size n = robots(in->size());
    const bool withRoot = m_fromRootModel and inRoot != nullptr;
    // one robot: the root pose sits at indexInput2 as in the reference; N robots: robot i uses root entry i
    const double *root = withRoot ? reinterpret_cast<const double *>(&(*inRoot)[n == 1 ? d_indexFromRoot.getValue() : 0]) : nullptr;
    assert(out->size() == n * static_cast<sofa::Size>(m_rbd.info.n_out));
    if (rbd_apply_host(m_rbd.ws, static_cast<int64_t>(n), reinterpret_cast<const double *>(in->data()), root,
                       reinterpret_cast<double *>(out->data())) != RBD_OK)
      fail("apply");

//@@ inl.applyJ.body
    assert(m_model);

    if (d_componentState.getValue() == sofa::core::objectmodel::ComponentState::Invalid)
      return;

    const sofa::Size n = static_cast<sofa::Size>(in->size() / (m_rbd.info.nv_joints > 0 ? m_rbd.info.nv_joints : 1));
    const bool withRoot = m_fromRootModel and inRoot != nullptr;
    // the root velocity is handed over verbatim, as the reference does (it is used as the free-flyer joint velocity)
    const double *rootVel = withRoot ? reinterpret_cast<const double *>(&(*inRoot)[n == 1 ? d_indexFromRoot.getValue() : 0]) : nullptr;
    // RBD_ERR_NO_APPLY when apply never ran: the Jacobians of the last apply are what applyJ uses, as in the reference
    if (rbd_applyJ_host(m_rbd.ws, static_cast<int64_t>(n > 0 ? n : 1), reinterpret_cast<const double *>(in->data()), rootVel,
                        reinterpret_cast<double *>(out->data())) != RBD_OK)
      fail("applyJ");

//@@ inl.applyJT.body
    assert(m_model);

    if (d_componentState.getValue() == sofa::core::objectmodel::ComponentState::Invalid)
      return;

    const sofa::Size n = static_cast<sofa::Size>(in->size() / (m_rbd.info.n_out > 0 ? m_rbd.info.n_out : 1));
    // both outputs are ACCUMULATED by the back end, as the reference does; the root wrench goes to entry 0 (robot i:
    // entry i), the reference's choice for the force side
    double *root = (m_fromRootModel and outRoot != nullptr) ? reinterpret_cast<double *>(&(*outRoot)[0]) : nullptr;
    if (rbd_applyJT_host(m_rbd.ws, static_cast<int64_t>(n > 0 ? n : 1), reinterpret_cast<const double *>(in->data()),
                         reinterpret_cast<double *>(out->data()), root) != RBD_OK)
      fail("applyJT");

//@@ inl.applyJT_matrix.body
    assert(m_model);

    if (d_componentState.getValue() == sofa::core::objectmodel::ComponentState::Invalid)
      return;

    // flatten the constraint matrix to CSR: a column index is a mapped-output index (robot r: r * n_out + k)
    const int nOut = m_rbd.info.n_out, nv = m_rbd.info.nv, nvJoints = m_rbd.info.nv_joints;
    m_rbdRows.clear();
    for (auto rowIt = in->begin(); rowIt != in->end(); ++rowIt)
    {
      int32_t robot = 0;
      for (auto colIt = rowIt.begin(); colIt != rowIt.end(); ++colIt)
      {
        const int32_t dof = static_cast<int32_t>(colIt.index());
        robot = dof / nOut; // a constraint line couples bodies of ONE robot
        const auto &w = colIt.val();
        const double w6[6] = {w.getVCenter()[0], w.getVCenter()[1], w.getVCenter()[2],
                              w.getVOrientation()[0], w.getVOrientation()[1], w.getVOrientation()[2]};
        m_rbdRows.addEntry(dof % nOut, w6);
      }
      m_rbdRows.endRow(static_cast<int32_t>(rowIt.index()), robot);
    }
    if (m_rbdRows.run(m_rbd.ws, nv) != RBD_OK)
    {
      fail("applyJT(MatrixDeriv)");
      return;
    }
    const bool withRoot = m_fromRootModel and outRoot != nullptr;
    for (int64_t r = 0; r < m_rbdRows.nrows(); ++r)
    {
      if (!m_rbdRows.kept(r)) // the reference drops lines whose torque is (numerically) null
        continue;
      const double *t = m_rbdRows.row(r, nv);
      const int32_t robot = m_rbdRows.row_instance[static_cast<size_t>(r)];
      auto outRowIt = out->writeLine(m_rbdRows.row_index[static_cast<size_t>(r)]);
      if (withRoot)
      {
        auto oRoot = outRoot->writeLine(m_rbdRows.row_index[static_cast<size_t>(r)]);
        oRoot.addCol(robot, Deriv_t<InRoot>(sofa::type::Vec3d(t[0], t[1], t[2]), sofa::type::Vec3d(t[3], t[4], t[5])));
        for (int i = 0; i < nvJoints; ++i)
          outRowIt.addCol(robot * nvJoints + i, Deriv_t<In>(t[6 + i]));
      }
      else
      {
        for (int i = 0; i < nv; ++i)
          outRowIt.addCol(robot * nv + i, Deriv_t<In>(t[i]));
      }
    }

//@@ end
