# reference translation units compiled in place (paths resolved through vpath in the Makefile)
REF_CPP := Array.cpp ProcessManager.cpp \
 Simulation.cpp FatalError.cpp Log.cpp TimeLogger.cpp Units.cpp SIUnits.cpp StellarUnits.cpp ExtragalacticUnits.cpp \
 Parallel.cpp ParallelFactory.cpp ParallelTarget.cpp ProcessAssigner.cpp IdenticalAssigner.cpp \
 SequentialAssigner.cpp StaggeredAssigner.cpp RootAssigner.cpp ProcessCommunicator.cpp PeerToPeerCommunicator.cpp \
 Random.cpp Position.cpp Direction.cpp StokesVector.cpp PhotonPackage.cpp DustGridPath.cpp \
 DustGrid.cpp BoxDustGrid.cpp CartesianDustGrid.cpp Mesh.cpp MoveableMesh.cpp LinMesh.cpp SymPowMesh.cpp PowMesh.cpp \
 SphereDustGrid.cpp Sphere1DDustGrid.cpp Sphere2DDustGrid.cpp CylinderDustGrid.cpp Cylinder2DDustGrid.cpp \
 TreeDustGrid.cpp ParticleTreeDustGrid.cpp OctTreeDustGrid.cpp BinTreeDustGrid.cpp TreeNode.cpp OctTreeNode.cpp BinTreeNode.cpp \
 BaryOctTreeNode.cpp BaryBinTreeNode.cpp TreeNodeBoxDensityCalculator.cpp TreeNodeSampleDensityCalculator.cpp \
 VoronoiDustGrid.cpp VoronoiMeshFile.cpp AdaptiveMesh.cpp AdaptiveMeshNode.cpp \
 AdaptiveMeshFile.cpp AdaptiveMeshDustGrid.cpp AdaptiveMeshDustDistribution.cpp BoxDustDistribution.cpp MeshDustComponent.cpp SpheroidalGeometryDecorator.cpp \
 MonteCarloSimulation.cpp OligoMonteCarloSimulation.cpp PanMonteCarloSimulation.cpp \
 DustSystem.cpp OligoDustSystem.cpp PanDustSystem.cpp DustSystemDensityCalculator.cpp DustSystemDepthCalculator.cpp \
 DustDistribution.cpp CompDustDistribution.cpp DustComp.cpp DustCompNormalization.cpp FaceOnDustCompNormalization.cpp DustMassDustCompNormalization.cpp \
 DustMix.cpp InterstellarDustMix.cpp DustLib.cpp AllCellsDustLib.cpp DustEmissivity.cpp GreyBodyDustEmissivity.cpp ISRF.cpp PlanckFunction.cpp \
 StellarSystem.cpp StellarComp.cpp GeometricStellarComp.cpp OligoStellarComp.cpp PanStellarComp.cpp \
 StellarCompNormalization.cpp BolLuminosityStellarCompNormalization.cpp SED.cpp StellarSED.cpp SunSED.cpp BlackBodySED.cpp \
 Geometry.cpp GenGeometry.cpp AxGeometry.cpp SepAxGeometry.cpp SpheGeometry.cpp ExpDiskGeometry.cpp SersicGeometry.cpp \
 SersicFunction.cpp SpecialFunctions.cpp SpiralStructureGeometryDecorator.cpp \
 WavelengthGrid.cpp OligoWavelengthGrid.cpp PanWavelengthGrid.cpp LogWavelengthGrid.cpp NestedLogWavelengthGrid.cpp \
 InstrumentSystem.cpp Instrument.cpp DistantInstrument.cpp SingleFrameInstrument.cpp FrameInstrument.cpp \
 SEDInstrument.cpp SimpleInstrument.cpp FullInstrument.cpp MultiFrameInstrument.cpp InstrumentFrame.cpp PerspectiveInstrument.cpp HomogeneousTransform.cpp
REF_CC := $(filter-out v_base_wl.cc,$(notdir $(wildcard $(REF)/Voro/*.cc)))
