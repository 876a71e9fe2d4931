package io.lubricant.consensus.raft.transport;
/** Compile-time stand-in: Membership only needs the nested marker type ID (RaftCluster.java:18). */
public interface RaftCluster { interface ID extends java.io.Serializable {} }
